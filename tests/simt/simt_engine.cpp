// simt_engine.cpp — TEST INFRASTRUCTURE (see simt_emu.h): the fiber scheduler behind ::simt::launch.
#include <sys/mman.h>

#include <mutex>
#include <thread>

#include "simt_emu.h"

namespace simt {

thread_local Fiber *g_cur = nullptr;

static const uint64_t g_seed = []() { const char *e = std::getenv("RDN_SIMT_SEED"); return e ? std::strtoull(e, nullptr, 10) : 0ull; }();
static thread_local uint64_t t_rng = 0;
static inline uint64_t rng_next() {  // xorshift64*
  t_rng ^= t_rng >> 12; t_rng ^= t_rng << 25; t_rng ^= t_rng >> 27;
  return t_rng * 0x2545F4914F6CDD1Dull;
}

static std::atomic<uint64_t> g_cost_warp[COST_REGIONS], g_cost_lane[COST_REGIONS];
// the lanes in `lanes` of warp `w` have met in a collective (or one lane leaves): charge the regions they passed since
static const bool g_cost_enabled = []() { const char *e = std::getenv("RDN_SIMT_COST"); return e && std::atoi(e) != 0; }();
static void cost_flush(Cta *cta, Warp &w, uint32_t lanes) {
  if (!g_cost_enabled) return;
  Fiber *base = &cta->fibers[static_cast<size_t>(&w - cta->warps.data()) * 32u];
  for (int r = 0; r < COST_REGIONS; ++r) {
    uint32_t mx = 0;
    uint64_t sum = 0;
    for (uint32_t l = 0; l < 32; ++l)
      if ((lanes >> l) & 1u) {
        uint32_t &c = base[l].cost[r];
        mx = c > mx ? c : mx;
        sum += c;
        c = 0;
      }
    if (mx) { g_cost_warp[r] += mx; g_cost_lane[r] += sum; }
  }
}

#if SIMT_FAST_SWITCH
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl simt_switch
    .hidden simt_switch
    .type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size simt_switch,.-simt_switch
)");
static inline void switch_context(Context &from, Context &to) { simt_switch(&from.sp, to.sp); }
#else
static inline void switch_context(Context &from, Context &to) { swapcontext(&from.uc, &to.uc); }
#endif

namespace {
constexpr size_t FIBER_STACK = 256 * 1024;

// fiber stacks of one OS thread, reused from CTA to CTA
struct StackPool {
  std::vector<char *> stacks;
  char *get(size_t i) {
    while (stacks.size() <= i) {
      void *p = mmap(nullptr, FIBER_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
      if (p == MAP_FAILED) { std::perror("simt: mmap"); std::abort(); }
      stacks.push_back(static_cast<char *>(p));
    }
    return stacks[i];
  }
  ~StackPool() { for (char *s : stacks) munmap(s, FIBER_STACK); }
};
thread_local StackPool t_stacks;

unsigned worker_count() {
  static const unsigned n = []() {
    const char *e = std::getenv("RDN_SIMT_THREADS");
    unsigned v = e ? static_cast<unsigned>(std::atoi(e)) : std::thread::hardware_concurrency();
    if (v < 1) v = 1;
    if (v > 16) v = 16;
    return v;
  }();
  return n;
}

void on_fiber_exit(Fiber *f) {
  Cta *cta = f->cta;
  Warp &w = *f->warp;
  f->done = true;
  cost_flush(cta, w, 1u << f->lane);
  w.exited |= 1u << f->lane;
  cta->n_done++;
  cta->progress++;
  // lanes that have left count as arrived (what the hardware does for exited threads)
  for (auto &kv : w.colls)
    if (kv.second.arrived) complete_if_ready(w, kv.second, kv.first);
  const uint32_t n_threads = static_cast<uint32_t>(cta->fibers.size());
  if (cta->sync_arrived && cta->sync_arrived + cta->n_done >= n_threads) {
    cta->sync_arrived = 0;
    cta->sync_gen++;
  }
}

void trampoline() {
  Fiber *f = g_cur;
  (*f->cta->body)();
  on_fiber_exit(f);
  switch_context(f->ctx, f->cta->sched);  // never resumed
  std::abort();
}

void prepare_fiber(Fiber &f, char *stack) {
#if SIMT_FAST_SWITCH
  // the frame simt_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into trampoline with the stack as after a call
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + FIBER_STACK) & ~static_cast<uintptr_t>(15);
  void **frame = reinterpret_cast<void **>(top - 32 - 8 * 6);  // six registers below the return slot, which sits on a 16-byte boundary
  for (int i = 0; i < 6; ++i) frame[i] = nullptr;
  frame[6] = reinterpret_cast<void *>(&trampoline);
  frame[7] = nullptr;  // the return address trampoline would see (it never returns)
  f.ctx.sp = frame;
#else
  getcontext(&f.ctx.uc);
  f.ctx.uc.uc_stack.ss_sp = stack;
  f.ctx.uc.uc_stack.ss_size = FIBER_STACK;
  f.ctx.uc.uc_link = nullptr;
  makecontext(&f.ctx.uc, trampoline, 0);
#endif
}

void run_cta(Cta &cta) {
  const uint32_t n_threads = cta.bdim.x * cta.bdim.y * cta.bdim.z;
  cta.fibers.assign(n_threads, Fiber{});
  cta.warps.assign((n_threads + 31u) / 32u, Warp{});
  for (uint32_t t = 0; t < n_threads; ++t) {
    Fiber &f = cta.fibers[t];
    f.tid = uint3{t % cta.bdim.x, (t / cta.bdim.x) % cta.bdim.y, t / (cta.bdim.x * cta.bdim.y)};
    f.lane = t & 31u;
    f.warp = &cta.warps[t >> 5];
    f.warp->exists |= 1u << f.lane;
    f.cta = &cta;
    prepare_fiber(f, t_stacks.get(t));
  }
  uint64_t idle_passes = 0;
  if (g_seed) t_rng = (g_seed * 0x9E3779B97F4A7C15ull) ^ (0xD1B54A32D192ED03ull * (1 + cta.bid.x + 65536ull * cta.bid.y)) | 1ull;
  while (cta.n_done < n_threads) {
    const uint64_t before = cta.progress;
    bool ran = false;
    cta.pass_id++;
    // randomised scheduling: a random rotation and a random odd stride (a permutation when the thread count is a power of two;
    // otherwise some threads are skipped in this pass and come up in a later one)
    const uint32_t rot = g_seed ? static_cast<uint32_t>(rng_next() % n_threads) : 0u;
    const uint32_t stride = g_seed ? static_cast<uint32_t>(rng_next() | 1u) : 1u;
    for (uint32_t k = 0; k < n_threads; ++k) {
      const uint32_t t = g_seed ? static_cast<uint32_t>((rot + static_cast<uint64_t>(k) * stride) % n_threads) : k;
      Fiber &f = cta.fibers[t];
      if (f.done) continue;
      if (f.wait_gen && *f.wait_gen == f.wait_val) continue;  // still blocked
      ran = true;
      g_cur = &f;
      switch_context(cta.sched, f.ctx);
    }
    g_cur = nullptr;
    if (!ran && cta.progress == before) {
      if (++idle_passes > 2) {
        std::fprintf(stderr, "simt: deadlock in CTA (%u,%u): every live thread waits in a collective that cannot complete\n",
                     cta.bid.x, cta.bid.y);
        for (uint32_t wi = 0; wi < cta.warps.size(); ++wi)
          for (auto &kv : cta.warps[wi].colls)
            if (kv.second.arrived)
              std::fprintf(stderr, "  warp %u mask %08x arrived %08x exited %08x\n", wi, kv.first, kv.second.arrived, cta.warps[wi].exited);
        std::abort();
      }
    } else {
      idle_passes = 0;
    }
  }
}
}  // namespace

void complete_if_ready(Warp &w, Coll &c, uint32_t mask) {
  const uint32_t need = mask & ~w.exited;
  if ((c.arrived & need) != need) return;
  std::memcpy(c.out, c.in, sizeof(c.in));
  c.out_mask = c.arrived & mask;
  c.out_pred = c.pred & c.out_mask;
  if (g_cur) cost_flush(g_cur->cta, w, c.out_mask);
  c.arrived = 0;
  c.pred = 0;
  c.gen++;
  if (g_cur) g_cur->cta->progress++;
}

void yield() {
  Fiber *f = g_cur;
  switch_context(f->ctx, f->cta->sched);
}

static void close_gather(Cta *cta, Warp &w, ActiveGather &g) {
  Fiber *base = &cta->fibers[static_cast<size_t>(&w - cta->warps.data()) * 32u];
  for (uint32_t l = 0; l < 32; ++l)
    if ((g.mask >> l) & 1u) base[l].active_result = g.mask;
  g.closed = true;
}

unsigned activemask_at(const void *site) {
  Fiber *f = g_cur;
  Cta *cta = f->cta;
  Warp &w = *f->warp;
  ActiveGather &g = w.gathers[site];
  if (!g.closed && g.pass != cta->pass_id) close_gather(cta, w, g);  // a group of an earlier pass that nobody has resumed from yet
  if (g.closed) { g.mask = 0; g.pass = cta->pass_id; g.closed = false; }
  g.mask |= 1u << f->lane;
  f->active_result = 0;
  f->wait_gen = nullptr;  // runnable: resumed in a later pass, when everyone who arrives in this one has
  cta->progress++;
  yield();
  if (f->active_result == 0) close_gather(cta, w, w.gathers[site]);
  return f->active_result;
}

void preempt() {
  if (!g_seed || !g_cur) return;
  if ((rng_next() & 3u) != 0) return;
  g_cur->wait_gen = nullptr;  // still runnable: picked up again in a later pass
  g_cur->cta->progress++;
  yield();
}

void syncthreads() {
  Fiber *f = g_cur;
  Cta *cta = f->cta;
  const uint32_t n_threads = static_cast<uint32_t>(cta->fibers.size());
  const uint64_t my_gen = cta->sync_gen;
  cta->sync_arrived++;
  if (cta->sync_arrived + cta->n_done >= n_threads) {
    cta->sync_arrived = 0;
    cta->sync_gen++;
    cta->progress++;
    return;
  }
  while (cta->sync_gen == my_gen) {
    f->wait_gen = &cta->sync_gen;
    f->wait_val = my_gen;
    yield();
  }
  f->wait_gen = nullptr;
}

void launch(dim3 grid, dim3 block, const std::function<void()> &body) {
  const uint64_t n_ctas = static_cast<uint64_t>(grid.x) * grid.y * grid.z;
  if (n_ctas == 0 || block.x * block.y * block.z == 0) return;
  std::atomic<uint64_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const uint64_t i = next.fetch_add(1);
      if (i >= n_ctas) break;
      Cta cta;
      cta.bid = uint3{static_cast<unsigned>(i % grid.x), static_cast<unsigned>((i / grid.x) % grid.y),
                      static_cast<unsigned>(i / (static_cast<uint64_t>(grid.x) * grid.y))};
      cta.bdim = block;
      cta.gdim = grid;
      cta.body = &body;
      run_cta(cta);
    }
  };
  const unsigned n_workers = static_cast<unsigned>(std::min<uint64_t>(worker_count(), n_ctas));
  // always on fresh threads: thread_local __shared__ storage and the fiber stacks belong to the worker
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < n_workers; ++w) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
}

}  // namespace simt

extern "C" void simt_cost_reset() {
  for (int r = 0; r < simt::COST_REGIONS; ++r) { simt::g_cost_warp[r] = 0; simt::g_cost_lane[r] = 0; }
}
extern "C" void simt_cost_read(uint64_t *warp_issues, uint64_t *lane_passes, int n) {
  for (int r = 0; r < n && r < simt::COST_REGIONS; ++r) { warp_issues[r] = simt::g_cost_warp[r]; lane_passes[r] = simt::g_cost_lane[r]; }
}
