// emu_runtime.cpp -- fiber scheduler, barriers and mbarrier emulation behind tools/emu/cuda_emu.h (TEST INFRASTRUCTURE).
#include "cuda_emu.h"

#include <sys/mman.h>
#include <map>
#include <vector>

// Context switch.  glibc's swapcontext saves and restores the signal mask with a system call on every switch, which
// dominates the run time of a barrier-heavy kernel; on x86-64 a switch only needs the callee-saved registers.
#if defined(__x86_64__)
extern "C" void b200emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl b200emu_switch
.type b200emu_switch,@function
b200emu_switch:
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
.size b200emu_switch,.-b200emu_switch
)");
#define B200EMU_ASM_SWITCH 1
#else
#include <ucontext.h>
#define B200EMU_ASM_SWITCH 0
#endif

namespace b200emu {

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Warp {
  unsigned exist = 0;     // lanes that exist in this block
  unsigned alive = 0;     // ... and have not returned yet
  unsigned arrived = 0;   // lanes waiting in warp_barrier
  unsigned released = 0;  // lanes whose barrier completed
  unsigned long long slots[32];
};

struct FiberState {
#if B200EMU_ASM_SWITCH
  void* sp = nullptr;
#else
  ucontext_t ctx;
#endif
  ThreadCtx tc;
  bool done = false;
  bool at_block_barrier = false;
  void* stack = nullptr;
};

struct MBar {
  unsigned count = 0;    // arrivals expected per phase
  unsigned pending = 0;  // arrivals still missing in the current phase
  long long tx = 0;      // bytes still expected in the current phase
  unsigned phase = 0;    // parity of the phase in progress
};

#if B200EMU_ASM_SWITCH
void* g_sched_sp = nullptr;
#else
ucontext_t g_sched;
#endif
std::vector<FiberState> g_fibers;
std::vector<Warp> g_warps;
int g_nthreads = 0;
int g_cur = -1;
int g_alive_threads = 0;
int g_block_arrived = 0;
unsigned long g_progress = 0;
const std::function<void()>* g_body = nullptr;
std::map<void*, MBar> g_mbars;
struct PendingCopy {  // an issued cp.async.bulk whose bytes have not been delivered yet
  void* dst;
  const void* src;
  unsigned bytes;
  void* bar;
};
std::vector<PendingCopy> g_pending;
alignas(128) unsigned char g_dyn_smem[256 * 1024];

void switch_out() {
  FiberState& f = g_fibers[g_cur];
#if B200EMU_ASM_SWITCH
  b200emu_switch(&f.sp, g_sched_sp);
#else
  swapcontext(&f.ctx, &g_sched);
#endif
}

void fiber_main() {
  (*g_body)();
  FiberState& f = g_fibers[g_cur];
  f.done = true;
  Warp& w = g_warps[f.tc.linear >> 5];
  w.alive &= ~(1u << (f.tc.linear & 31));
  --g_alive_threads;
  ++g_progress;
  switch_out();  // never resumed: the scheduler skips finished fibers
  abort();
}

void fiber_prepare(FiberState& f) {
#if B200EMU_ASM_SWITCH
  // stack image b200emu_switch pops: six callee-saved registers, then `ret` into fiber_main with the stack pointer
  // where a call instruction would have left it (16-byte aligned before the pushed return address)
  uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;              // return address of fiber_main (it never returns)
  *--sp = (void*)&fiber_main;   // popped by `ret`
  for (int r = 0; r < 6; ++r) *--sp = nullptr;
  f.sp = sp;
#else
  getcontext(&f.ctx);
  f.ctx.uc_stack.ss_sp = f.stack;
  f.ctx.uc_stack.ss_size = kStackBytes;
  f.ctx.uc_link = &g_sched;
  makecontext(&f.ctx, fiber_main, 0);
#endif
}

void switch_in(FiberState& f) {
#if B200EMU_ASM_SWITCH
  b200emu_switch(&g_sched_sp, f.sp);
#else
  swapcontext(&g_sched, &f.ctx);
#endif
}

void try_release_block() {
  if (g_block_arrived > 0 && g_block_arrived == g_alive_threads) {
    for (int t = 0; t < g_nthreads; ++t) g_fibers[t].at_block_barrier = false;
    g_block_arrived = 0;
    ++g_progress;
  }
}

}  // namespace

// B200EMU_GUARD=1: every "device" allocation ends (rounded up to 16 bytes, the widest vector access) right before an
// inaccessible page, and is preceded by one, so that a kernel reading or writing past either end of an exact-size array
// faults here instead of passing silently -- what a cudaMalloc granule or torch's caching allocator would hide on a GPU.
static int guard_mode() {
  static int g = -1;
  if (g < 0) {
    const char* e = getenv("B200EMU_GUARD");
    g = (e && atoi(e) > 0) ? 1 : 0;
  }
  return g;
}
static std::map<void*, std::pair<void*, size_t>> g_guarded;  // user pointer -> (mapping, mapped bytes)

// ---------------------------------------------------------------------------------------------- schedule order
static int g_order = -1;  // 0 forward, 1 reverse, 2 random
static unsigned long long g_rng = 0x9E3779B97F4A7C15ull;
static std::vector<int> g_perm;
static void order_init() {
  if (g_order >= 0) return;
  g_order = 0;
  if (const char* e = getenv("B200EMU_ORDER")) {
    if (!strncmp(e, "reverse", 7)) g_order = 1;
    if (!strncmp(e, "random", 6)) {
      g_order = 2;
      if (e[6] == ':') g_rng ^= strtoull(e + 7, nullptr, 10) * 0xD1B54A32D192ED03ull;
    }
  }
}
static int pass_order(int i, int n) {
  if (g_order == 0) return i;
  if (g_order == 1) return n - 1 - i;
  if (i == 0) {  // new pass: new permutation (Fisher-Yates with a xorshift generator)
    g_perm.resize(n);
    for (int k = 0; k < n; ++k) g_perm[k] = k;
    for (int k = n - 1; k > 0; --k) {
      g_rng ^= g_rng << 13;
      g_rng ^= g_rng >> 7;
      g_rng ^= g_rng << 17;
      std::swap(g_perm[k], g_perm[(int)(g_rng % (unsigned)(k + 1))]);
    }
  }
  return g_perm[i];
}

// Guarded block for test inputs (tests/emu_lib.py): `bytes` of payload ending `slack` bytes before an inaccessible page
// (slack = 0: the first byte past the array faults; the array then starts wherever that puts it, aligned to `align`).
static void* guarded_block(size_t bytes, size_t align) {
  const size_t page = 4096;
  const size_t payload = bytes ? bytes : 1;
  const size_t body = (payload + align + page - 1) / page * page;
  const size_t total = body + 2 * page;
  char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == (char*)MAP_FAILED) return nullptr;
  mprotect(base, page, PROT_NONE);
  mprotect(base + page + body, page, PROT_NONE);
  char* end = base + page + body;
  char* user = (char*)(((uintptr_t)(end - payload)) & ~(uintptr_t)(align - 1));
  memset(base + page, 0xA5, body);
  g_guarded[user] = {base, total};
  return user;
}

// live "device" allocations (both modes): b200emu_live_allocations() lets a test assert that destroying a plan gives back
// everything the library allocated for it -- a leak a long-running solver would pay for in HBM
static std::map<void*, size_t> g_live;
static void* device_alloc_impl(size_t bytes);
void* device_alloc(size_t bytes) {
  void* p = device_alloc_impl(bytes);
  if (p) g_live[p] = bytes;
  return p;
}
static void* device_alloc_impl(size_t bytes) {
  if (!guard_mode()) {
    void* p = nullptr;
    if (posix_memalign(&p, 256, bytes ? bytes : 1) != 0) return nullptr;
    return p;
  }
  const size_t page = 4096;
  const size_t payload = ((bytes ? bytes : 1) + 15) & ~(size_t)15;
  const size_t body = (payload + page - 1) / page * page;
  const size_t total = body + 2 * page;
  char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == (char*)MAP_FAILED) return nullptr;
  mprotect(base, page, PROT_NONE);
  mprotect(base + page + body, page, PROT_NONE);
  char* user = base + page + body - payload;
  memset(base + page, 0xA5, body);  // uninitialised device memory is not zero
  g_guarded[user] = {base, total};
  return user;
}
void device_free(void* p) {
  if (!p) return;
  if (g_live.erase(p) == 0) {
    fprintf(stderr, "b200emu: cudaFree of a pointer cudaMalloc did not return (or freed twice)\n");
    abort();
  }
  if (!guard_mode()) {
    free(p);
    return;
  }
  auto it = g_guarded.find(p);
  if (it == g_guarded.end()) {
    fprintf(stderr, "b200emu: cudaFree of a pointer cudaMalloc did not return\n");
    abort();
  }
  munmap(it->second.first, it->second.second);
  g_guarded.erase(it);
}

ThreadCtx* cur() { return &g_fibers[g_cur].tc; }
void* dyn_smem() { return g_dyn_smem; }
void note_progress() { ++g_progress; }
unsigned long long* warp_slots() { return g_warps[g_fibers[g_cur].tc.linear >> 5].slots; }
unsigned warp_alive() { return g_warps[g_fibers[g_cur].tc.linear >> 5].alive; }

void yield_blocked() { switch_out(); }

void block_barrier() {
  FiberState& f = g_fibers[g_cur];
  f.at_block_barrier = true;
  ++g_block_arrived;
  ++g_progress;
  try_release_block();
  while (f.at_block_barrier) {
    switch_out();
    try_release_block();  // threads may have exited meanwhile
  }
}

void warp_barrier(unsigned mask) {
  FiberState& f = g_fibers[g_cur];
  Warp& w = g_warps[f.tc.linear >> 5];
  const unsigned bit = 1u << (f.tc.linear & 31);
  if (!(mask & bit)) {
    fprintf(stderr, "b200emu: lane %d calls a warp collective with mask %08x that does not name it\n", f.tc.linear & 31, mask);
    abort();
  }
  w.arrived |= bit;
  ++g_progress;
  for (;;) {
    if (w.released & bit) {
      w.released &= ~bit;
      return;
    }
    const unsigned need = mask & w.alive;
    if ((w.arrived & need) == need) {
      w.arrived &= ~need;
      w.released |= need;
      ++g_progress;
      continue;  // picks up its own release
    }
    switch_out();
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > kMaxThreads || smem > sizeof(g_dyn_smem)) {
    fprintf(stderr, "b200emu: bad launch configuration (%d threads, %zu bytes of dynamic shared memory)\n", nthreads, smem);
    abort();
  }
  if (g_cur >= 0) {
    fprintf(stderr, "b200emu: nested launch\n");
    abort();
  }
  if ((int)g_fibers.size() < nthreads) {
    const size_t old = g_fibers.size();
    g_fibers.resize(nthreads);
    for (size_t t = old; t < g_fibers.size(); ++t) {
      g_fibers[t].stack = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (g_fibers[t].stack == MAP_FAILED) {
        perror("b200emu: mmap");
        abort();
      }
    }
  }
  order_init();
  g_body = &body;
  g_nthreads = nthreads;
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_warps.assign(nwarps, Warp());
        g_mbars.clear();
        g_pending.clear();
        g_alive_threads = nthreads;
        g_block_arrived = 0;
        for (int t = 0; t < nthreads; ++t) {
          FiberState& f = g_fibers[t];
          f.done = false;
          f.at_block_barrier = false;
          f.tc.linear = t;
          f.tc.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          f.tc.bid = uint3{bx, by, bz};
          f.tc.bdim = block;
          f.tc.gdim = grid;
          g_warps[t >> 5].exist |= 1u << (t & 31);
          g_warps[t >> 5].alive |= 1u << (t & 31);
          fiber_prepare(f);
        }
        // round-robin until every fiber has returned; a full window of passes without progress is a deadlock.  The
        // order of a pass is a knob (B200EMU_ORDER=forward|reverse|random[:seed]): code that is correct only because
        // lower threads happen to run first (a missing __syncwarp / __syncthreads) fails under the other orders.  Under
        // `random` every warp of a block also gets its own speed (it runs in one pass out of 1..4), so that a producer
        // warp can run ahead of slow consumers, or the other way round: a missing "slot is free" wait, which equal speeds
        // never expose, corrupts the result here.
        unsigned char period[kMaxThreads / 32];
        for (int w = 0; w < nwarps; ++w) {
          period[w] = 1;
          if (g_order == 2) {
            g_rng ^= g_rng << 13;
            g_rng ^= g_rng >> 7;
            g_rng ^= g_rng << 17;
            period[w] = (unsigned char)(1 + (g_rng >> 20) % 4);
          }
        }
        constexpr unsigned kWindow = 12;  // lcm of the periods: every warp has run at least once
        unsigned pass = 0;
        unsigned long window_start = g_progress;
        while (g_alive_threads > 0) {
          for (int i = 0; i < nthreads; ++i) {
            const int t = pass_order(i, nthreads);
            if (g_fibers[t].done || pass % period[t >> 5] != 0) continue;
            g_cur = t;
            switch_in(g_fibers[t]);
          }
          g_cur = -1;
          if (++pass % kWindow == 0) {
            if (g_progress == window_start && g_alive_threads > 0) {
              fprintf(stderr, "b200emu: deadlock in block (%u,%u,%u): %d threads alive, %d at the block barrier\n", bx, by, bz,
                      g_alive_threads, g_block_arrived);
              for (int w = 0; w < nwarps; ++w)
                fprintf(stderr, "  warp %d: alive %08x arrived %08x released %08x\n", w, g_warps[w].alive, g_warps[w].arrived,
                        g_warps[w].released);
              abort();
            }
            window_start = g_progress;
          }
        }
        g_cur = -1;
      }
  g_body = nullptr;
}

// ---------------------------------------------------------------------------------------------- mbarrier
void mbar_init(void* bar, unsigned count) {
  MBar& b = g_mbars[bar];
  b.count = b.pending = count;
  b.tx = 0;
  b.phase = 0;
  ++g_progress;
}
static void mbar_check(MBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = b.count;
    ++g_progress;
  }
}
void mbar_arrive(void* bar) {
  MBar& b = g_mbars[bar];
  if (b.pending == 0) {
    fprintf(stderr, "b200emu: mbarrier over-arrival\n");
    abort();
  }
  --b.pending;
  ++g_progress;
  mbar_check(b);
}
void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
  MBar& b = g_mbars[bar];
  b.tx += bytes;
  if (b.pending == 0) {
    fprintf(stderr, "b200emu: mbarrier over-arrival\n");
    abort();
  }
  --b.pending;
  ++g_progress;
  mbar_check(b);
}
void mbar_complete_tx(void* bar, unsigned bytes) {
  MBar& b = g_mbars[bar];
  b.tx -= bytes;
  ++g_progress;
  mbar_check(b);
}
// cp.async.bulk global -> shared.  The copy is ASYNCHRONOUS on the hardware: nothing may read the destination before
// its mbarrier completes the phase.  To make code that breaks this rule fail here too, the destination is filled with
// 0xFF at issue time and the bytes are delivered only when some thread polls that barrier.

void bulk_copy_async(void* smem_dst, const void* gsrc, unsigned bytes, void* bar) {
  if ((((uintptr_t)smem_dst | (uintptr_t)gsrc | bytes) & 15u) != 0) {
    fprintf(stderr, "b200emu: cp.async.bulk with unaligned operands (dst %p src %p bytes %u)\n", smem_dst, gsrc, bytes);
    abort();
  }
  if ((unsigned char*)smem_dst < g_dyn_smem || (unsigned char*)smem_dst + bytes > g_dyn_smem + sizeof(g_dyn_smem)) {
    fprintf(stderr, "b200emu: cp.async.bulk destination outside dynamic shared memory\n");
    abort();
  }
  memset(smem_dst, 0xFF, bytes);
  g_pending.push_back({smem_dst, gsrc, bytes, bar});
  ++g_progress;
}

static void deliver_pending(void* bar) {
  size_t keep = 0;
  for (size_t i = 0; i < g_pending.size(); ++i) {
    PendingCopy c = g_pending[i];
    if (c.bar == bar) {
      memcpy(c.dst, c.src, c.bytes);
      mbar_complete_tx(bar, c.bytes);
    } else {
      g_pending[keep++] = c;
    }
  }
  g_pending.resize(keep);
}

bool mbar_test_wait(void* bar, unsigned parity) {
  // true once the phase with the given parity has completed, i.e. the phase in progress has the other parity
  deliver_pending(bar);
  MBar& b = g_mbars[bar];
  if (b.phase != parity) return true;
  switch_out();
  return g_mbars[bar].phase != parity;
}

}  // namespace b200emu

// C entry points for tests that want their INPUT arrays guarded too (numpy views over these blocks, tests/emu_lib.py)
extern "C" {
__attribute__((visibility("default"))) void* b200emu_guarded_alloc(size_t bytes, size_t align) {
  if (align == 0 || (align & (align - 1)) != 0) return nullptr;
  return b200emu::guarded_block(bytes, align);
}
__attribute__((visibility("default"))) long long b200emu_live_allocations(long long* bytes) {
  long long b = 0;
  for (auto& kv : b200emu::g_live) b += (long long)kv.second;
  if (bytes) *bytes = b;
  return (long long)b200emu::g_live.size();
}
__attribute__((visibility("default"))) void b200emu_guarded_free(void* p) {
  auto it = b200emu::g_guarded.find(p);
  if (it == b200emu::g_guarded.end()) abort();
  munmap(it->second.first, it->second.second);
  b200emu::g_guarded.erase(it);
}
}
