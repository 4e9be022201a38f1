// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
// A serial interpreter for the CUDA kernels of fdtd_b200/csrc: the .cu file is compiled as
// plain C++ with -DFDTD_EMU, every <<<grid, block>>> launch becomes a loop that runs the
// kernel body once per (block, thread) with threadIdx / blockIdx set.  This works because
// the streaming kernels have no inter-thread communication (no shared memory, shuffles or barriers).
// Kernels that do use shared memory and __syncthreads() (the temporally fused E+H kernels) are run by
// emu::launch_coop: the threads of a block become cooperative fibers (ucontext) on the one host thread,
// __syncthreads() hands control to the next fiber, `__shared__` is a function-local static.
// It lets the CPU test-suite (`pytest -m "not gpu"`) execute the kernel logic, the host
// layer and the 2-rank halo exchange (gloo) in a container without a GPU.  The product
// never loads this build: fdtd_b200 only opens libfdtd_b200.so (nvcc, sm_100a).
#pragma once
#include <stddef.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

template <typename F>
inline void launch(dim3 grid, dim3 block, F&& body) {
  t_gridDim = grid;
  t_blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              t_blockIdx = dim3(bx, by, bz);
              t_threadIdx = dim3(tx, ty, tz);
              body();
            }
}
}  // namespace emu

#include <string.h>
#include <ucontext.h>

#include <type_traits>
#include <vector>

namespace emu {
// ---- cooperative launch: one fiber per thread of a block, for kernels with barriers ---------------------
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  dim3 tid;
  unsigned long shfl_gen = 0;
};
inline thread_local ucontext_t t_sched;
inline thread_local Fiber* t_cur = nullptr;
inline thread_local void (*t_entry)(void*) = nullptr;
inline thread_local void* t_entry_arg = nullptr;

inline void fiber_exit();
inline void fiber_main() {
  t_entry(t_entry_arg);
  t_cur->done = true;
  fiber_exit();
  swapcontext(&t_cur->ctx, &t_sched);
}


// hand control to the next fiber of the block (every spin loop below does)
inline void yield() { swapcontext(&t_cur->ctx, &t_sched); }
// __syncthreads(): a counting barrier over the live threads of the block (fibers are not in lockstep: the shuffle
// emulation below lets the lanes of a warp rendezvous on their own)
inline thread_local unsigned t_bar_arrived = 0, t_bar_live = 0;
inline thread_local unsigned long t_bar_gen = 0;
inline void sync_threads() {
  const unsigned long gen = t_bar_gen;
  if (++t_bar_arrived >= t_bar_live) {
    t_bar_arrived = 0;
    ++t_bar_gen;
    return;
  }
  while (t_bar_gen == gen) yield();
}

// a thread that has exited no longer counts in the block barrier
inline void fiber_exit() {
  --t_bar_live;
  if (t_bar_live > 0 && t_bar_arrived >= t_bar_live) {
    t_bar_arrived = 0;
    ++t_bar_gen;
  }
}

// ---- TMA + mbarrier emulation (the fused E+H kernel's staging): a bulk tensor copy is DEFERRED to the first wait on
// its barrier -- the latest moment the hardware may complete it -- and fills what lies outside the tensor with zeros
struct TmaCopy {
  void* dst;
  const char* base;
  size_t esize;
  long n0, n1, n2;
  long c0, c1, c2;
  int b0, b1;
  const void* bar;
};
inline thread_local std::vector<TmaCopy> t_tma;
inline void mbar_init(const void* bar) {
  for (size_t k = 0; k < t_tma.size();)
    if (t_tma[k].bar == bar) t_tma.erase(t_tma.begin() + k); else ++k;
}
inline void tma_load_3d(void* dst, const void* base, size_t esize, long n0, long n1, long n2, long c0, long c1, long c2,
                        int b0, int b1, const void* bar) {
  // from the moment a copy is issued its destination may change at any time: poison it (NaN bit patterns), so that a
  // thread still reading the stage's previous contents -- a missing "stage is free" rendezvous -- shows in the results
  memset(dst, 0xff, (size_t)b0 * b1 * esize);
  t_tma.push_back({dst, (const char*)base, esize, n0, n1, n2, c0, c1, c2, b0, b1, bar});
}
// cp.async.bulk (1-D): `bytes` contiguous bytes, deferred like the tensor copies (n0 < 0 marks it, c0 = bytes)
inline void bulk_load(void* dst, const void* src, size_t bytes, const void* bar) {
  memset(dst, 0xff, bytes);
  t_tma.push_back({dst, (const char*)src, 1, -1, 0, 0, (long)bytes, 0, 0, 0, 0, bar});
}
inline void mbar_wait(const void* bar) {
  for (size_t k = 0; k < t_tma.size();) {
    const TmaCopy& c = t_tma[k];
    if (c.bar != bar) { ++k; continue; }
    if (c.n0 < 0) {
      memcpy(c.dst, c.base, (size_t)c.c0);
      t_tma.erase(t_tma.begin() + k);
      continue;
    }
    char* out = (char*)c.dst;
    for (int y = 0; y < c.b1; ++y)
      for (int z = 0; z < c.b0; ++z, out += c.esize) {
        const long gz = c.c0 + z, gy = c.c1 + y, gx = c.c2;
        if (gz < 0 || gz >= c.n0 || gy < 0 || gy >= c.n1 || gx < 0 || gx >= c.n2) memset(out, 0, c.esize);
        else memcpy(out, c.base + ((gx * c.n1 + gy) * c.n0 + gz) * c.esize, c.esize);
      }
    t_tma.erase(t_tma.begin() + k);
  }
}

// __shfl_down_sync(full mask, v, 1) of a one-dimensional block whose rows are warps: the lanes of a warp rendezvous
// (warps are NOT in lockstep with each other once the block barrier is split), deposit their values in the slot of
// this shuffle's generation and read their upper neighbour's (the last lane of a warp gets its own back)
inline thread_local double t_shfl[2][1024];
inline thread_local unsigned long t_shfl_arrived[32];   // per warp: deposits so far
template <typename V>
inline V shfl_down1(V v) {
  const unsigned t = t_cur->tid.x, w = t >> 5;
  const unsigned lanes = t_blockDim.x - (w << 5) < 32u ? t_blockDim.x - (w << 5) : 32u;
  const unsigned long gen = t_cur->shfl_gen++;
  t_shfl[gen & 1][t] = (double)v;
  ++t_shfl_arrived[w];
  while (t_shfl_arrived[w] < (gen + 1) * lanes) yield();   // (a lane cannot be a whole generation ahead)
  return ((t & 31u) != 31u && t + 1 < t_blockDim.x) ? (V)t_shfl[gen & 1][t + 1] : v;
}

template <typename F>
inline void launch_coop(dim3 grid, dim3 block, F&& body) {
  t_gridDim = grid;
  t_blockDim = block;
  const unsigned n = block.x * block.y * block.z;
  static thread_local std::vector<Fiber> fibers;
  if (fibers.size() < n) fibers.resize(n);
  for (Fiber& f : fibers)
    if (f.stack.empty()) f.stack.resize(256 << 10);
  using Body = typename std::remove_reference<F>::type;
  t_entry = [](void* p) { (*static_cast<Body*>(p))(); };
  t_entry_arg = const_cast<void*>(static_cast<const void*>(&body));
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        unsigned t = 0;
        for (unsigned long& a : t_shfl_arrived) a = 0;
        t_bar_arrived = 0;
        t_bar_live = n;
        t_bar_gen = 0;
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
              Fiber& f = fibers[t];
              f.done = false;
              f.shfl_gen = 0;
              f.tid = dim3(tx, ty, tz);
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = f.stack.data();
              f.ctx.uc_stack.ss_size = f.stack.size();
              f.ctx.uc_link = nullptr;
              makecontext(&f.ctx, fiber_main, 0);
            }
        for (bool alive = true; alive;) {
          alive = false;
          for (unsigned k = 0; k < n; ++k) {
            Fiber& f = fibers[k];
            if (f.done) continue;
            t_blockIdx = dim3(bx, by, bz);
            t_threadIdx = f.tid;
            t_cur = &f;
            swapcontext(&t_sched, &f.ctx);
            alive = alive || !f.done;
          }
        }
      }
}
}  // namespace emu

#define __syncthreads() emu::sync_threads()

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
