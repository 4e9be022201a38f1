// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
// A serial interpreter for the CUDA kernels of fdtd_b200/csrc: the .cu file is compiled as
// plain C++ with -DFDTD_EMU, every <<<grid, block>>> launch becomes a loop that runs the
// kernel body once per (block, thread) with threadIdx / blockIdx set.  This works because
// the kernels have no inter-thread communication (no shared memory, shuffles or barriers).
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

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
