// simt_selftest.cpp — TEST INFRASTRUCTURE: known answers for the emulator's own primitives (called through ctypes by
// tests/test_simt_emulation.py).  Each function returns 0 on success or the line of the first failed check.
#include "simt_emu.h"

#define CHECK(cond) do { if (!(cond)) { int expected = 0; __atomic_compare_exchange_n(fail, &expected, __LINE__, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); } } while (0)

namespace {

void k_collectives(int *fail, unsigned *out) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // full-mask vote and shuffles
  const unsigned even = __ballot_sync(0xFFFFFFFFu, (lane & 1u) == 0);
  CHECK(even == 0x55555555u);
  CHECK(__shfl_sync(0xFFFFFFFFu, lane * 3u + warp, 7) == 21u + warp);
  CHECK(__shfl_down_sync(0xFFFFFFFFu, lane, 4) == (lane + 4 < 32 ? lane + 4 : lane));
  CHECK(__shfl_up_sync(0xFFFFFFFFu, lane, 2) == (lane >= 2 ? lane - 2 : lane));
  CHECK(__all_sync(0xFFFFFFFFu, 1) == 1 && __any_sync(0xFFFFFFFFu, lane == 31) == 1 && __all_sync(0xFFFFFFFFu, lane != 5) == 0);
  unsigned long long wide = (static_cast<unsigned long long>(lane) << 40) | 5u;
  CHECK(__shfl_sync(0xFFFFFFFFu, wide, 9) == ((9ull << 40) | 5u));
  // two groups of a diverged warp use different masks at the same time; the groups run a different number of rounds
  const unsigned mine = (lane < 12) ? 0x00000FFFu : 0xFFFFF000u;
  const int rounds = (lane < 12) ? 5 : 2;
  unsigned acc = 0;
  for (int r = 0; r < rounds; ++r) {
    acc += __popc(__ballot_sync(mine, true));
    __syncwarp(mine);
  }
  CHECK(acc == (lane < 12 ? 5u * 12u : 2u * 20u));
  __syncwarp();
  // block barrier + shared memory: every warp publishes, everyone reads all of them
  __shared__ unsigned s_sum[8];
  if (lane == 0) s_sum[warp] = warp + 1;
  __syncthreads();
  unsigned total = 0;
  for (unsigned w = 0; w < blockDim.x / 32; ++w) total += s_sum[w];
  CHECK(total == (blockDim.x / 32) * (blockDim.x / 32 + 1) / 2);
  __syncthreads();
  // lanes that have left count as arrived
  if (lane >= 16) { atomicAdd(out, 1u); return; }
  CHECK(__ballot_sync(0xFFFFFFFFu, true) == 0x0000FFFFu);
  atomicAdd(out, 1u);
  atomicMax(out + 1, blockIdx.x * 1000u + threadIdx.x);
  atomicMin(out + 2, blockIdx.x * 1000u + threadIdx.x);
}

// a deliberate race: read-modify-write of one word with a preemption point in the middle
void k_race(unsigned *word) {
  const unsigned v = *reinterpret_cast<volatile unsigned *>(word);
  __threadfence();  // (a preemption point of the randomised scheduler)
  *reinterpret_cast<volatile unsigned *>(word) = v + 1u;
}

}  // namespace

extern "C" int simt_selftest_collectives() {
  int fail = 0;
  unsigned out[3] = {0, 0, 0xFFFFFFFFu};
  ::simt::launch(dim3(3), dim3(128), [&]() { k_collectives(&fail, out); });
  if (fail) return fail;
  if (out[0] != 3u * 128u) return -1;
  if (out[1] != 2000u + 15u + 96u || out[2] != 0u) return -2;   // largest / smallest id among the lanes that stayed
  return 0;
}

// number of increments that survive out of 64 racing threads of one CTA: 64 only when no two threads interleave
extern "C" unsigned simt_selftest_race() {
  unsigned word = 0;
  ::simt::launch(dim3(1), dim3(64), [&]() { k_race(&word); });
  return word;
}

namespace {
// warp-aggregated append (the pattern of traverse.cu's enqueue_rewalk): every participating lane must get a distinct slot
void k_aggregate(unsigned *count, unsigned *slots, unsigned *groups_seen) {
  const unsigned lane = threadIdx.x & 31u;
  if ((lane % 3u) == 0u) return;  // a third of the lanes do not take part
  const unsigned peers = __activemask();
  const int leader = __ffs(peers) - 1;
  unsigned base = 0;
  if (static_cast<int>(lane) == leader) base = atomicAdd(count, static_cast<unsigned>(__popc(peers)));
  base = __shfl_sync(peers, base, leader);
  const unsigned q = base + __popc(peers & ((1u << lane) - 1u));
  slots[q] = threadIdx.x + 1u;
  if (__popc(peers) > 1) atomicAdd(groups_seen, 1u);
}
}  // namespace

// returns the number of lanes that found themselves in a group of more than one, or -1 when two lanes got the same slot
extern "C" int simt_selftest_aggregate() {
  unsigned count = 0, groups = 0;
  unsigned slots[128] = {};
  ::simt::launch(dim3(1), dim3(128), [&]() { k_aggregate(&count, slots, &groups); });
  unsigned expected = 0;
  for (unsigned t = 0; t < 128; ++t) expected += (t & 31u) % 3u != 0u;
  if (count != expected) return -1;
  unsigned long long seen_lo = 0, seen_hi = 0;
  for (unsigned q = 0; q < count; ++q) {
    if (slots[q] == 0 || slots[q] > 128) return -1;
    unsigned long long &word = slots[q] <= 64 ? seen_lo : seen_hi;
    const unsigned long long bit = 1ull << ((slots[q] - 1u) & 63u);
    if (word & bit) return -1;
    word |= bit;
  }
  return static_cast<int>(groups);
}
