// simt_emu.h — TEST INFRASTRUCTURE: runs the CUDA kernels of rendiation_b200/csrc on the CPU, thread for thread.
//
// This is not a CPU implementation of the product and the package never loads anything built from it.  It exists so that the
// kernel SOURCE (the same .cu files nvcc compiles for sm_100a) can be executed in the CPU test run: tests/simt/build_emu.py
// compiles the sources with g++, this header force-included, into tests/simt/_build/librdn_rt_emu.so, and
// tests/test_simt_emulation.py drives that library through the same C-ABI and compares it with the oracle.  What it checks is
// the kernels' LOGIC (traversal order, tie handling, queue protocol, refill, warp votes, scans); what it cannot check is anything
// that depends on the hardware (memory model, scheduling, timing, PTX paths replaced under RDN_SIMT_EMU).
//
// Execution model: a CTA runs on one OS thread as blockDim cooperatively scheduled fibers (ucontext); CTAs of a grid are dealt
// to a few OS threads.  A fiber runs until it reaches a warp collective (__ballot_sync, __shfl*_sync, __syncwarp, ...) or
// __syncthreads and is resumed when every lane named by the mask (that has not exited) has arrived.  Collectives on different
// masks may be pending in a warp at the same time (diverged lanes).  __activemask() returns the lanes that reach the call site
// in the same scheduler pass (a legal answer on hardware with independent thread scheduling, and one that exercises
// warp-aggregated code with real groups).  __shared__ variables are thread_local statics (one CTA per OS thread
// at a time).  Global-memory atomics are host atomics, so CTAs on different OS threads interact as CTAs on different SMs do.
// Arithmetic: compiled with -ffp-contract=off, SSE f32 — bit-identical to the -fmad=false device code except NaN payloads.
#pragma once
#include <cuda_runtime.h>
#include <ucontext.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>
#include <vector>

#undef __shared__
#define __shared__ static thread_local
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__

using std::isinf;
using std::isnan;

namespace simt {

struct Cta;
struct Coll {  // one pending collective per (warp, mask)
  uint32_t arrived = 0, pred = 0;
  uint64_t gen = 0;
  uint64_t in[32];
  uint64_t out[32];
  uint32_t out_pred = 0, out_mask = 0;
};
struct ActiveGather {  // lanes that reached one __activemask() call site during the same scheduler pass
  uint32_t mask = 0;
  uint64_t pass = ~0ull;
  bool closed = true;
};
struct Warp {
  uint32_t exists = 0, exited = 0;
  std::unordered_map<uint32_t, Coll> colls;  // node based: references stay valid
  std::unordered_map<const void *, ActiveGather> gathers;
};
constexpr int COST_REGIONS = 32;
// Context switch: on x86-64 a hand-written one (callee-saved registers and the stack pointer; glibc's swapcontext makes a
// sigprocmask system call per switch, which was half of the emulation's run time), elsewhere ucontext.
#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_FAST_SWITCH 1
struct Context {
  void *sp = nullptr;
};
#else
#define SIMT_FAST_SWITCH 0
struct Context {
  ucontext_t uc;
};
#endif
struct Fiber {
  uint32_t cost[COST_REGIONS] = {};  // RDN_SIMT_COST: passes through each marked region since the lane's last collective
  Context ctx;
  uint3 tid;
  uint32_t lane = 0;
  Warp *warp = nullptr;
  Cta *cta = nullptr;
  bool done = false;
  const volatile uint64_t *wait_gen = nullptr;  // blocked while *wait_gen == wait_val
  uint64_t wait_val = 0;
  uint32_t active_result = 0;                   // __activemask(): filled when the lane's gather closes
};
struct Cta {
  uint3 bid;
  dim3 bdim, gdim;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  uint32_t sync_arrived = 0, n_done = 0;
  uint64_t sync_gen = 0;
  uint64_t progress = 0;
  uint64_t pass_id = 0;                         // scheduler passes over the CTA's threads so far
  Context sched;
  const std::function<void()> *body = nullptr;
};

extern thread_local Fiber *g_cur;
void yield();                                                              // back to the CTA scheduler
void launch(dim3 grid, dim3 block, const std::function<void()> &body);     // synchronous: returns when the grid has finished
void complete_if_ready(Warp &w, Coll &c, uint32_t mask);

inline const Coll &collective(uint32_t mask, uint64_t value, bool pred) {
  Fiber *f = g_cur;
  Warp &w = *f->warp;
  mask &= w.exists;
  if (!((mask >> f->lane) & 1u)) {
    std::fprintf(stderr, "simt: lane %u calls a collective whose mask %08x does not name it\n", f->lane, mask);
    std::abort();
  }
  Coll &c = w.colls[mask];
  c.in[f->lane] = value;
  if (pred) c.pred |= 1u << f->lane;
  c.arrived |= 1u << f->lane;
  const uint64_t my_gen = c.gen;
  complete_if_ready(w, c, mask);
  while (c.gen == my_gen) {
    f->wait_gen = &c.gen;
    f->wait_val = my_gen;
    yield();
  }
  f->wait_gen = nullptr;
  return c;
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}

void syncthreads();

// RDN_SIMT_COST: a SIMT issue model.  Kernels mark code regions (RDN_COST(region) in the sources); when a collective
// completes, every region costs the warp max-over-its-lanes passes since their previous collective (lanes that run the same
// region run it in lockstep, different regions serialise) — the number of times the warp ISSUES that region.  Multiplied by the
// region's static SASS instruction count (tools/sass_regions.py) this predicts the kernel's warp-level instruction count.
inline void cost_mark(int region) { g_cur->cost[region]++; }

// RDN_SIMT_SEED=n (n != 0): randomised scheduling.  The CTA scheduler visits its threads in a random order each pass and a thread
// may be preempted at every atomic, fence, volatile / .cg load and __nanosleep, so lanes of a warp and warps of a CTA interleave
// differently from run to run of different seeds — results that depend on the interleaving (a missing fence or barrier whose
// effect the round-robin order happens to hide) show up as parity failures.  Independent thread scheduling in miniature.
void preempt();

}  // namespace simt
extern "C" void simt_cost_reset();
extern "C" void simt_cost_read(uint64_t *warp_issues, uint64_t *lane_passes, int n);

// ---- built-in variables (objects, not macros: cudaLaunchConfig_t has members called gridDim / blockDim)
namespace simt {
struct Comp {
  uint8_t var, comp;
  operator unsigned() const {
    const Fiber *f = g_cur;
    const unsigned v[4][3] = {{f->tid.x, f->tid.y, f->tid.z},
                              {f->cta->bid.x, f->cta->bid.y, f->cta->bid.z},
                              {f->cta->bdim.x, f->cta->bdim.y, f->cta->bdim.z},
                              {f->cta->gdim.x, f->cta->gdim.y, f->cta->gdim.z}};
    return v[var][comp];
  }
};
struct Builtin {
  Comp x, y, z;
};
}  // namespace simt
static constexpr ::simt::Builtin threadIdx{{0, 0}, {0, 1}, {0, 2}}, blockIdx{{1, 0}, {1, 1}, {1, 2}}, blockDim{{2, 0}, {2, 1}, {2, 2}},
    gridDim{{3, 0}, {3, 1}, {3, 2}};

// ---- warp / CTA collectives
inline unsigned __ballot_sync(unsigned mask, int pred) { return ::simt::collective(mask, 0, pred != 0).out_pred; }
inline int __any_sync(unsigned mask, int pred) { return ::simt::collective(mask, 0, pred != 0).out_pred != 0; }
inline int __all_sync(unsigned mask, int pred) {
  const ::simt::Coll &c = ::simt::collective(mask, 0, pred != 0);
  return c.out_pred == c.out_mask;
}
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { ::simt::collective(mask, 0, false); }
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  (void)width;
  const ::simt::Coll &c = ::simt::collective(mask, ::simt::to_bits(v), false);
  src &= 31;
  return ((c.out_mask >> src) & 1u) ? ::simt::from_bits<T>(c.out[src]) : v;
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  (void)width;
  const unsigned lane = ::simt::g_cur->lane;
  const ::simt::Coll &c = ::simt::collective(mask, ::simt::to_bits(v), false);
  const unsigned src = lane + delta;
  return (src < 32u && ((c.out_mask >> src) & 1u)) ? ::simt::from_bits<T>(c.out[src]) : v;
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  (void)width;
  const unsigned lane = ::simt::g_cur->lane;
  const ::simt::Coll &c = ::simt::collective(mask, ::simt::to_bits(v), false);
  return (delta <= lane && ((c.out_mask >> (lane - delta)) & 1u)) ? ::simt::from_bits<T>(c.out[lane - delta]) : v;
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  (void)width;
  const unsigned lane = ::simt::g_cur->lane;
  const ::simt::Coll &c = ::simt::collective(mask, ::simt::to_bits(v), false);
  const unsigned src = (lane ^ static_cast<unsigned>(lane_mask)) & 31u;
  return ((c.out_mask >> src) & 1u) ? ::simt::from_bits<T>(c.out[src]) : v;
}
// __activemask(): the lanes of the warp that reach the same call site within one scheduler pass form a converged group (all of
// them then continue from that site, so a __shfl_sync over the returned mask completes).  Under RDN_SIMT_SEED the groups vary.
namespace simt { unsigned activemask_at(const void *site); }
__attribute__((noinline)) inline unsigned __activemask() { return ::simt::activemask_at(__builtin_return_address(0)); }
inline unsigned __fns(unsigned mask, unsigned base, int offset) {  // offset-th set bit at or above `base` (offset >= 1), else 0xFFFFFFFF
  for (unsigned b = base; b < 32; ++b)
    if (((mask >> b) & 1u) && --offset == 0) return b;
  return 0xFFFFFFFFu;
}
inline void __syncthreads() { ::simt::syncthreads(); }
inline void __threadfence() { ::simt::preempt(); __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { ::simt::preempt(); __builtin_ia32_pause(); }

// ---- atomics on global memory
template <class T, class U>
inline T atomicAdd(T *p, U v) {
  ::simt::preempt();
  if constexpr (std::is_floating_point<T>::value) {
    T old = *reinterpret_cast<volatile T *>(p);
    for (;;) {
      const T want = old + static_cast<T>(v);
      T expected = old;
      if (__atomic_compare_exchange(p, &expected, const_cast<T *>(&want), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return old;
      old = expected;
    }
  } else {
    return __atomic_fetch_add(p, static_cast<T>(v), __ATOMIC_SEQ_CST);
  }
}
template <class T, class U>
inline T atomicMin(T *p, U v) {
  ::simt::preempt();
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (static_cast<T>(v) < old && !__atomic_compare_exchange_n(p, &old, static_cast<T>(v), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T, class U>
inline T atomicMax(T *p, U v) {
  ::simt::preempt();
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (static_cast<T>(v) > old && !__atomic_compare_exchange_n(p, &old, static_cast<T>(v), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T, class U, class V>
inline T atomicCAS(T *p, U compare, V value) {
  ::simt::preempt();
  T expected = static_cast<T>(compare);
  __atomic_compare_exchange_n(p, &expected, static_cast<T>(value), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;
}

// ---- loads, bit casts, integer helpers
template <class T>
inline T __ldg(const T *p) { return *p; }
template <class T>
inline T __ldcg(const T *p) {
  ::simt::preempt();
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  T v;
  std::memcpy(&v, const_cast<const T *>(p), sizeof(T));
  return v;
}
inline unsigned __float_as_uint(float f) { return ::simt::from_bits<unsigned>(::simt::to_bits(f)); }
inline int __float_as_int(float f) { return ::simt::from_bits<int>(::simt::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return ::simt::from_bits<float>(::simt::to_bits(u)); }
inline float __int_as_float(int i) { return ::simt::from_bits<float>(::simt::to_bits(i)); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs(static_cast<int>(x)); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }

// cudaLaunchKernelEx / kernel<<<...>>>: build_emu.py rewrites every `k<<<grid, block, ...>>>(args);` into
// `::simt::launch(grid, block, [&]() { k(args); });`; the extended launch goes through this overload.
template <class... Params, class... Args>
inline cudaError_t simt_launch_ex(const cudaLaunchConfig_t *cfg, void (*kernel)(Params...), Args &&...args) {
  ::simt::launch(cfg->gridDim, cfg->blockDim, [&]() { kernel(args...); });
  return cudaSuccess;
}
#define cudaLaunchKernelEx simt_launch_ex
