// fdtd_b200.cu -- C ABI (include/fdtd_b200.h) over the sm_100a Yee kernels.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -shared -Xcompiler -fPIC -Iinclude fdtd_b200/csrc/fdtd_b200.cu -o fdtd_b200/libfdtd_b200.so
// (see __graft_entry__.build()).  The same file compiles as plain C++ with -DFDTD_EMU
// against tests/emu/cuda_emu.h: a serial thread-by-thread interpreter of the kernels (cooperative fibers for the
// few kernels with barriers) used
// ONLY by the CPU test-suite to exercise the kernel logic where there is no GPU.  The
// product library never contains that path.
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "fdtd_b200.h"

#ifdef FDTD_EMU
#include "cuda_emu.h"
#define FDTD_DEV inline
#define FDTD_LAUNCH(kern, grid, block, stream, ...) \
  emu::launch(grid, block, [&]() { kern(__VA_ARGS__); })
// kernels with shared memory and barriers: the threads of a block run as cooperative fibers
#define FDTD_LAUNCH_SYNC(kern, grid, block, stream, ...) \
  emu::launch_coop(grid, block, [&]() { kern(__VA_ARGS__); })
#define FDTD_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) \
  (emu::t_tma.clear(), emu::launch_coop(grid, block, [&]() { kern(__VA_ARGS__); }))
template <typename T>
inline void fdtd_atomic_add(T* p, T v) { *p = *p + v; }
#include <cmath>
template <typename T>
inline bool fdtd_signbit(T v) { return std::signbit(v); }
#else
#include <cuda.h>            // CUtensorMap and its enums only: the encoder is resolved at run time (no libcuda link)
#include <cuda_runtime.h>
#define FDTD_DEV __device__ __forceinline__
#define FDTD_LAUNCH(kern, grid, block, stream, ...) \
  kern<<<grid, block, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define FDTD_LAUNCH_SYNC FDTD_LAUNCH
#define FDTD_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) \
  kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__)
template <typename T>
__device__ __forceinline__ void fdtd_atomic_add(T* p, T v) { atomicAdd(p, v); }
template <typename T>
__device__ __forceinline__ bool fdtd_signbit(T v) { return signbit(v); }
#endif

#include "yee_kernels.cuh"
#include "yee_fused_eh.cuh"

namespace {

using fdtd::i64;

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FDTD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return FDTD_OK;
}

struct Geometry {
  int vec;          // cells per thread
  int lanes_z;      // threads along z in a block (power of two)
  int lanes_shift;
  int rows;         // y rows per block (power of two)
  int tile_y, tile_z;
};

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

#ifndef FDTD_FUSE_MAX_CELLS
#define FDTD_FUSE_MAX_CELLS (1LL << 25)
#endif
#ifndef FDTD_MAX_LANES_Z
#define FDTD_MAX_LANES_Z 32   // threads of a block along z (x VEC cells each); the rest of the 256 go to y rows
#endif
#ifndef FDTD_MAX_VEC_F32
#define FDTD_MAX_VEC_F32 4
#endif

Geometry geometry(int dtype, int Ny, int Nz) {
  Geometry g;
  if (dtype == FDTD_F32 || dtype == FDTD_F32X)
    g.vec = (Nz % 4 == 0 && FDTD_MAX_VEC_F32 >= 4) ? 4 : ((Nz % 2 == 0 && FDTD_MAX_VEC_F32 >= 2) ? 2 : 1);
  else
    g.vec = (Nz % 2 == 0) ? 2 : 1;
  int nvz = (Nz + g.vec - 1) / g.vec;
  g.lanes_z = pow2_ceil(nvz);
  // rows of up to 128 vectors: 16 lanes x 16 rows (fewer warps touch the z-PML ends of a row: +6 % at 512^3 f32
  // and 256^3 f64); longer rows: 32 lanes x 8 rows (-1 % otherwise at 1024^3)            profiles/r1_tune8
  const int max_lanes = nvz <= 128 ? (FDTD_MAX_LANES_Z < 16 ? FDTD_MAX_LANES_Z : 16) : FDTD_MAX_LANES_Z;
  if (g.lanes_z > max_lanes) g.lanes_z = max_lanes;
  g.lanes_shift = 0;
  while ((1 << g.lanes_shift) < g.lanes_z) ++g.lanes_shift;
  g.rows = FDTD_BLOCK_THREADS / g.lanes_z;
  int ny2 = pow2_ceil(Ny);
  if (g.rows > ny2) g.rows = ny2;
  g.tile_y = g.rows;
  g.tile_z = g.lanes_z * g.vec;
  return g;
}

// planes marched per block.  Two costs pull in opposite directions (fitted to profiles/r1_tune3.txt and
// r1_tune4_slabs.txt on B200): every chunk re-reads two carried planes (~0.32/chunk of a launch) and the
// last, partially filled wave of blocks runs the memory system below capacity (~0.8/waves, 444 resident
// blocks).  Pick the power of two in [4, 32] that minimises their sum.
int default_x_chunk(const Geometry& g, int nx, int Ny, int Nz, int blocks_per_sm = 3) {
  const double tiles = (double)((Ny + g.tile_y - 1) / g.tile_y) * ((Nz + g.tile_z - 1) / g.tile_z);
  const double resident = 148.0 * blocks_per_sm;
  int best = 4;
  double best_cost = 1e30;
  // grids too small to fill the GPU even with 4-plane chunks are latency bound: the march is serial, so the
  // shortest chunk (most blocks) wins
  // (161x97x1 quick-start grid: 15.2 / 20.3 / 27.1 / 45.5 us per step with 1 / 2 / 4 / 8 planes per block)
  if (((nx + 3) / 4) * tiles <= 2 * resident) {
    for (int chunk = 1; chunk < 4; chunk *= 2)
      if (((nx + chunk - 1) / chunk) * tiles <= 4 * resident) return chunk;
    return 4;
  }
  for (int chunk = 4; chunk <= 32; chunk *= 2) {
    const double waves = ((nx + chunk - 1) / chunk) * tiles / resident;
    const double cost = 0.32 / chunk + 0.8 / waves;
    if (cost < best_cost) {
      best_cost = cost;
      best = chunk;
    }
  }
  return best;
}

// z slabs: psi rows start at z = lo rounded down to a multiple of 4 and are padded to a multiple of 4,
// so that a thread's VEC cells are one aligned vector of its psi row
int z_slab_lo(const fdtd_slab& S) { return S.lo & ~3; }
int z_slab_row(const fdtd_slab& S) { return ((S.lo + S.thickness - z_slab_lo(S)) + 3) & ~3; }

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int validate(const fdtd_desc* d) {
  if (!d) return fail(FDTD_ERR_ARG, "null descriptor");
  if (d->abi_version != FDTD_ABI_VERSION)
    return fail(FDTD_ERR_ARG, "descriptor ABI %d != library ABI %d", d->abi_version, FDTD_ABI_VERSION);
  if (d->dtype != FDTD_F32 && d->dtype != FDTD_F64 && d->dtype != FDTD_F32X)
    return fail(FDTD_ERR_ARG, "bad dtype %d", d->dtype);
  if (d->Nx < 1 || d->Ny < 1 || d->Nz < 1) return fail(FDTD_ERR_ARG, "bad extents");
  if (d->plane != (int64_t)d->Ny * d->Nz) return fail(FDTD_ERR_ARG, "plane != Ny*Nz");
  if (d->x_offset < 0 || d->x_offset + d->Nx > d->Nx_global)
    return fail(FDTD_ERR_ARG, "slab [%d,%d) outside global Nx=%d", d->x_offset, d->x_offset + d->Nx, d->Nx_global);
  Geometry g = geometry(d->dtype, d->Ny, d->Nz);
  const size_t w = d->dtype == FDTD_F64 ? 8 : 4;     // state: fields, psi, rings
  const size_t wa = d->dtype == FDTD_F32 ? 4 : 8;    // coefficients: material arrays, tables, profiles, waveforms
  for (int c = 0; c < 3; ++c) {
    if (!d->E[c] || !d->H[c]) return fail(FDTD_ERR_ARG, "null field pointer");
    if (!aligned(d->E[c], w * g.vec) || !aligned(d->H[c], w * g.vec))
      return fail(FDTD_ERR_ARG, "field pointer not aligned to %zu bytes", w * g.vec);
    const void* opt[6] = {d->inv_eps[c], d->inv_eps_grid[c], d->absorb[c], d->inv_mu[c], d->inv_eps2[c],
                          d->absorb2[c]};
    for (int n = 0; n < 6; ++n)
      if (opt[n] && !aligned(opt[n], wa * g.vec)) return fail(FDTD_ERR_ARG, "material pointer misaligned");
  }
  bool any_e = d->inv_eps[0] || d->inv_eps[1] || d->inv_eps[2];
  bool all_e = d->inv_eps[0] && d->inv_eps[1] && d->inv_eps[2];
  bool any_m = d->inv_mu[0] || d->inv_mu[1] || d->inv_mu[2];
  bool all_m = d->inv_mu[0] && d->inv_mu[1] && d->inv_mu[2];
  if (any_e != all_e || any_m != all_m) return fail(FDTD_ERR_ARG, "material arrays must be given for all 3 components");
  if (d->tile_class && (d->tile_y != g.tile_y || d->tile_z != g.tile_z))
    return fail(FDTD_ERR_ARG, "tile_class laid out for %dx%d tiles, kernels use %dx%d", d->tile_y, d->tile_z,
                g.tile_y, g.tile_z);
  if (d->n_slabs < 0 || d->n_slabs > FDTD_MAX_SLABS) return fail(FDTD_ERR_ARG, "n_slabs");
  for (int s = 0; s < d->n_slabs; ++s) {
    const fdtd_slab& S = d->slabs[s];
    int n_axis = S.axis == 0 ? d->Nx_global : (S.axis == 1 ? d->Ny : d->Nz);
    if (S.axis < 0 || S.axis > 2 || S.thickness < 1 || S.lo < 0 || S.lo + S.thickness > n_axis)
      return fail(FDTD_ERR_ARG, "slab %d geometry", s);
    if (S.x0 < 0 || S.x1 > d->Nx || S.x0 > S.x1) return fail(FDTD_ERR_ARG, "slab %d x-range", s);
    i64 want = S.axis == 0 ? (i64)(S.x1 - S.x0) * d->plane
                           : (S.axis == 1 ? (i64)d->Nx * S.thickness * d->Nz : (i64)d->Nx * d->Ny * z_slab_row(S));
    if (S.psi_count != want) return fail(FDTD_ERR_ARG, "slab %d psi_count %lld != %lld", s, (long long)S.psi_count, want);
    if (want > 0 && (!S.psi_E || !S.psi_H || !S.bE || !S.cE || !S.bH || !S.cH))
      return fail(FDTD_ERR_ARG, "slab %d null pointer", s);
    if (want > 0 && (!aligned(S.psi_E, w * g.vec) || !aligned(S.psi_H, w * g.vec)))
      return fail(FDTD_ERR_ARG, "slab %d psi misaligned", s);
  }
  if (d->n_post < 0 || d->n_post > FDTD_MAX_POST) return fail(FDTD_ERR_ARG, "n_post");
  for (int n = 0; n < d->n_post; ++n) {
    if (d->post_kind[n] == FDTD_POST_PERIODIC) {
      if (d->post_arg[n] < 0 || d->post_arg[n] > 2) return fail(FDTD_ERR_ARG, "periodic axis");
      if (d->post_arg[n] == 0 && d->Nx != d->Nx_global)
        return fail(FDTD_ERR_UNSUPPORTED, "periodic x boundary on an x-sharded grid");
    } else if (d->post_kind[n] == FDTD_POST_PML_ADD) {
      if (d->post_arg[n] < 0 || d->post_arg[n] >= d->n_slabs) return fail(FDTD_ERR_ARG, "post slab index");
    } else {
      return fail(FDTD_ERR_ARG, "post kind");
    }
  }
  if (d->n_sources < 0 || (d->n_sources > 0 && !d->sources)) return fail(FDTD_ERR_ARG, "n_sources");
  for (int n = 0; n < d->n_sources; ++n) {
    const fdtd_source& S = d->sources[n];
    if (S.field < 0 || S.field > 1 || S.comp < 0 || S.comp > 2) return fail(FDTD_ERR_ARG, "source %d field/comp", n);
    if (S.kind == FDTD_SRC_POINTS) {
      if (S.n < 0 || (S.n > 0 && (!S.idx || !S.profile))) return fail(FDTD_ERR_ARG, "source %d points", n);
    } else if (S.kind == FDTD_SRC_BOX) {
      if (S.box[0] < 0 || S.box[1] > d->Nx || S.box[2] < 0 || S.box[3] > d->Ny || S.box[4] < 0 || S.box[5] > d->Nz)
        return fail(FDTD_ERR_ARG, "source %d box", n);
    } else if (S.kind == FDTD_SRC_FEEDBACK) {
      if (S.n < 0 || S.n > 1 || (S.n == 1 && (!S.feedback || !S.profile || S.spacing <= 0 || S.field != 0)))
        return fail(FDTD_ERR_ARG, "source %d feedback", n);
      if (S.n == 1 && (S.box[0] < 0 || S.box[0] >= d->Nx)) return fail(FDTD_ERR_ARG, "source %d feedback cell", n);
    } else {
      return fail(FDTD_ERR_ARG, "source %d kind", n);
    }
    if (!S.wave || S.wave_len < 1) return fail(FDTD_ERR_ARG, "source %d wave table", n);
  }
  if (d->n_detectors < 0 || (d->n_detectors > 0 && !d->detectors)) return fail(FDTD_ERR_ARG, "n_detectors");
  for (int n = 0; n < d->n_detectors; ++n) {
    const fdtd_detector& D = d->detectors[n];
    if (D.kind != FDTD_DET_FIELD && D.kind != FDTD_DET_CURRENT) return fail(FDTD_ERR_ARG, "detector %d kind", n);
    if (D.n < 0 || (D.n > 0 && (!D.idx || !D.pos || !D.ring_H || D.capacity < 1)))
      return fail(FDTD_ERR_ARG, "detector %d", n);
    if (D.n > 0 && D.kind == FDTD_DET_FIELD && !D.ring_E) return fail(FDTD_ERR_ARG, "detector %d ring_E", n);
    if (D.n > 0 && D.kind == FDTD_DET_CURRENT && (!D.last || D.spacing <= 0))
      return fail(FDTD_ERR_ARG, "detector %d current", n);
  }
  if (d->n_deep < 0 || (d->n_deep > 0 && !d->deep)) return fail(FDTD_ERR_ARG, "n_deep");
  for (int n = 0; n < d->n_deep; ++n) {
    const fdtd_deep_object& O = d->deep[n];
    if (O.kind < FDTD_OBJ_PLAIN || O.kind > FDTD_OBJ_ABSORB) return fail(FDTD_ERR_ARG, "deep object %d kind", n);
    if (O.box[0] < 0 || O.box[1] > d->Nx || O.box[2] < 0 || O.box[3] > d->Ny || O.box[4] < 0 || O.box[5] > d->Nz)
      return fail(FDTD_ERR_ARG, "deep object %d box", n);
    const bool empty = O.box[0] >= O.box[1] || O.box[2] >= O.box[3] || O.box[4] >= O.box[5];
    if (!empty && (!O.mask || !O.inv[0] || !O.inv[1] || !O.inv[2] ||
                   (O.kind == FDTD_OBJ_ABSORB && (!O.absorb[0] || !O.absorb[1] || !O.absorb[2]))))
      return fail(FDTD_ERR_ARG, "deep object %d null pointer", n);
  }
  if (d->use_graphs && !d->dyn) return fail(FDTD_ERR_ARG, "use_graphs needs the dyn scratch");
  if (d->x_wrap < 0 || d->x_wrap > d->n_post + 1) return fail(FDTD_ERR_ARG, "x_wrap outside the post op list");
  if (d->x_wrap && d->Nx == d->Nx_global)
    return fail(FDTD_ERR_ARG, "x_wrap is for x-sharded slabs; an unsharded grid lists the periodic x boundary as a post op");
  return FDTD_OK;
}

// periodic copies on a one-cell axis are the identity (E[0] = E[-1]); anything else between the field
// update and the sources forbids folding sources / detectors into the half-step kernel
bool post_is_fused(const fdtd_desc* d) {
  // folding pays where a step is launch-bound; on large slabs the separate source / detector kernels
  // cost < 0.1 % of a step while the folded code costs ~1 % of the streaming kernel (profiles/r1_tune6)
  // (objects applied by their own kernels after the fused one come BEFORE the sources, fdtd/grid.py:285-295)
  if (d->fuse_post == 0 || d->x_wrap || d->n_deep > 0) return false;
  if (d->fuse_post < 0 && (int64_t)d->Nx * d->Ny * d->Nz > FDTD_FUSE_MAX_CELLS) return false;
  for (int n = 0; n < d->n_post; ++n) {
    if (d->post_kind[n] == FDTD_POST_PML_ADD) return false;
    int axis = d->post_arg[n];
    int N = axis == 0 ? d->Nx : (axis == 1 ? d->Ny : d->Nz);
    if (N >= 2) return false;
  }
  int ns[2] = {0, 0};
  for (int n = 0; n < d->n_sources; ++n) {
    if (d->sources[n].kind == FDTD_SRC_FEEDBACK) return false;
    ns[d->sources[n].field]++;
  }
  for (int n = 0; n < d->n_detectors; ++n)
    if (d->detectors[n].kind != FDTD_DET_FIELD) return false;
  if (ns[0] > FDTD_FUSED_MAX || ns[1] > FDTD_FUSED_MAX) return false;
  int nd = 0;
  for (int n = 0; n < d->n_detectors; ++n) nd += d->detectors[n].n > 0;
  return nd <= FDTD_FUSED_MAX;
}

template <typename T>
T rounded_product(double a, double b) {
  // sc * inverse material as the reference computes it: both operands in the grid dtype
  return (T)((T)a * (T)b);
}

template <typename T, bool IS_E, typename A = T>
fdtd::SlabK<T, A> slab_k(const fdtd_slab& S) {
  fdtd::SlabK<T, A> k;
  k.axis = S.axis;
  k.lo = S.lo;
  k.t = S.thickness;
  k.fused = S.fused;
  k.x0 = S.x0;
  k.x1 = S.x1;
  k.count = S.psi_count;
  k.lo_al = z_slab_lo(S);
  k.tp = z_slab_row(S);
  k.psi = (T*)(IS_E ? S.psi_E : S.psi_H);
  k.b = (const A*)(IS_E ? S.bE : S.bH);
  k.c = (const A*)(IS_E ? S.cE : S.cH);
  return k;
}

// explicit in / out / curl-source buffers of a launch (the last H plane of a temporally fused step on an x-sharded
// slab: H from one buffer of the ping-pong pair into the other, curls from the new E); no folded sources / detectors
struct Buffers {
  void* const* Fin;
  void* const* Fout;
  void* const* G;
};

// graph_step >= 0: the launch is being captured as step `graph_step` of a replayable chunk; waveform
// index and ring slot are then graph_step + the bases in d->dyn
template <typename T, bool IS_E, typename A = T>
int launch_halfstep_run(const fdtd_desc* d, int x_begin, int x_end, int64_t q, int64_t slot, void* stream,
                        int64_t graph_step, void* push_y, void* push_z, bool plain, const Buffers* buf = nullptr);

#ifndef FDTD_EMU
// The runs of one half-step touch disjoint x-planes of the updated field and only read the other one: they are
// independent, so they go to a few side streams (forked from / joined into the caller's stream with events) and fill
// the GPU together instead of each ending in its own partial wave of blocks.
struct RunStreams {
  static const int N = 3;
  cudaStream_t side[N] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[N] = {nullptr, nullptr, nullptr};
  bool ready = false;
  bool ensure() {
    if (ready) return true;
    if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return false;
    for (int n = 0; n < N; ++n)
      if (cudaStreamCreateWithFlags(&side[n], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&join[n], cudaEventDisableTiming) != cudaSuccess)
        return false;
    ready = true;
    return true;
  }
};
thread_local RunStreams g_runs;
#endif

#ifndef FDTD_PLAIN_RUN_MIN
#ifdef FDTD_EMU
#define FDTD_PLAIN_RUN_MIN 2   // (CPU tests: small scenes must take the split path too)
#else
#define FDTD_PLAIN_RUN_MIN 8   // shorter runs of object-free planes are not worth a launch of their own
#endif
#endif

// the planes [x_begin, x_end) of a half-step: one launch, or -- when the descriptor says which x-planes carry
// material classes at all -- one launch per run of planes with / without them, the latter with the material-free
// instantiation of the kernel
template <typename T, bool IS_E, typename A = T>
int launch_halfstep(const fdtd_desc* d, int x_begin, int x_end, int64_t q, int64_t slot, void* stream,
                    int64_t graph_step = -1, void* push_y = nullptr, void* push_z = nullptr) {
  if (x_begin < 0 || x_end > d->Nx || x_begin > x_end) return fail(FDTD_ERR_ARG, "plane range [%d,%d)", x_begin, x_end);
  if (x_begin == x_end) return FDTD_OK;
  if (!d->plane_class || !d->tile_class || post_is_fused(d))
    return launch_halfstep_run<T, IS_E, A>(d, x_begin, x_end, q, slot, stream, graph_step, push_y, push_z, false);
  // the class bits this half-step reads: everything but the mu^-1 bit for E, only that bit for H
  const unsigned char bits = IS_E ? (unsigned char)~FDTD_CLS_VARY_H : (unsigned char)FDTD_CLS_VARY_H;
  auto is_plain = [&](int x) { return (d->plane_class[x] & bits) == 0; };
  const int push_plane = IS_E ? 0 : d->Nx - 1;
  int i = x_begin;
  int n_run = 0;
#ifndef FDTD_EMU
  unsigned used = 0;
  const bool concurrent = g_runs.ensure() && cudaEventRecord(g_runs.fork, (cudaStream_t)stream) == cudaSuccess;
#endif
  while (i < x_end) {
    // the next run: planes of one kind; a short plain run is absorbed by the material run around it
    bool plain = is_plain(i);
    int j = i + 1;
    while (j < x_end && is_plain(j) == plain) ++j;
    if (plain && j - i < FDTD_PLAIN_RUN_MIN && !(i == x_begin && j == x_end)) plain = false;
    if (!plain) {
      for (;;) {     // extend over material planes and short plain runs
        while (j < x_end && !is_plain(j)) ++j;
        int k = j;
        while (k < x_end && is_plain(k)) ++k;
        if (k == j || k - j >= FDTD_PLAIN_RUN_MIN) break;
        j = k;
      }
    }
    const bool has = push_y && push_plane >= i && push_plane < j;
    void* run_stream = stream;
#ifndef FDTD_EMU
    // the first run stays on the caller's stream, the others rotate over the side streams
    if (concurrent && n_run > 0) {
      const int k = (n_run - 1) % RunStreams::N;
      if (!(used & (1u << k)) && cudaStreamWaitEvent(g_runs.side[k], g_runs.fork, 0) != cudaSuccess)
        return fail(FDTD_ERR_CUDA, "cudaStreamWaitEvent failed");
      used |= 1u << k;
      run_stream = g_runs.side[k];
    }
#endif
    int rc = launch_halfstep_run<T, IS_E, A>(d, i, j, q, slot, run_stream, graph_step, has ? push_y : nullptr,
                                             has ? push_z : nullptr, plain);
    if (rc) return rc;
    ++n_run;
    i = j;
  }
#ifndef FDTD_EMU
  for (int k = 0; k < RunStreams::N; ++k) {
    if (!(used & (1u << k))) continue;
    if (cudaEventRecord(g_runs.join[k], g_runs.side[k]) != cudaSuccess ||
        cudaStreamWaitEvent((cudaStream_t)stream, g_runs.join[k], 0) != cudaSuccess)
      return fail(FDTD_ERR_CUDA, "joining the run streams failed");
  }
#endif
  return FDTD_OK;
}

// plain: every tile of these planes is homogeneous -> no class map, no material arrays
template <typename T, bool IS_E, typename A>
int launch_halfstep_run(const fdtd_desc* d, int x_begin, int x_end, int64_t q, int64_t slot, void* stream,
                        int64_t graph_step, void* push_y, void* push_z, bool plain, const Buffers* buf) {
  Geometry g = geometry(d->dtype, d->Ny, d->Nz);
  fdtd::HalfStepParams<T, A> P;
  memset(&P, 0, sizeof(P));
  P.Nx = d->Nx;
  P.Ny = d->Ny;
  P.Nz = d->Nz;
  P.x_offset = d->x_offset;
  P.Nx_global = d->Nx_global;
  P.x_begin = x_begin;
  P.x_end = x_end;
  P.x_chunk = d->x_chunk > 0 ? d->x_chunk
                             : default_x_chunk(g, x_end - x_begin, d->Ny, d->Nz,
                                               (sizeof(A) > sizeof(T) || (!plain && d->tile_class)) ? 2 : 3);
  P.lanes_z = g.lanes_z;
  P.lanes_shift = g.lanes_shift;
  P.rows = g.rows;
  P.plane = d->plane;
  P.sc = (A)d->courant;
  P.y_begin = 0;
  P.y_end = d->Ny;
  P.z_begin = 0;
  P.z_end = d->Nz;
  for (int c = 0; c < 3; ++c) {
    P.F[c] = (T*)(IS_E ? d->E[c] : d->H[c]);
    P.Fo[c] = P.F[c];
    P.G[c] = (const T*)(IS_E ? d->H[c] : d->E[c]);
    P.bg_c[c] = rounded_product<A>(d->courant, IS_E ? d->bg_inv_eps[c] : d->bg_inv_mu[c]);
    P.bg_inv[c] = (A)(IS_E ? d->bg_inv_eps[c] : d->bg_inv_mu[c]);
    P.inv[c] = (const A*)(IS_E ? d->inv_eps[c] : d->inv_mu[c]);
    P.inv_grid[c] = IS_E ? (const A*)d->inv_eps_grid[c] : nullptr;
    P.inv2[c] = IS_E ? (const A*)d->inv_eps2[c] : nullptr;
    P.absorb[c] = IS_E ? (const A*)d->absorb[c] : nullptr;
    P.absorb2[c] = IS_E ? (const A*)d->absorb2[c] : nullptr;
    if (buf) {
      P.F[c] = (T*)buf->Fin[c];
      P.Fo[c] = (T*)buf->Fout[c];
      P.G[c] = (const T*)buf->G[c];
    }
  }
  // a class map is only meaningful with the arrays it refers to
  P.cls = plain ? nullptr : d->tile_class;
  P.cls_vary = (P.inv[0] != nullptr && !plain) ? (IS_E ? FDTD_CLS_VARY_E : FDTD_CLS_VARY_H) : 0;
  if (P.inv[0] != nullptr && d->tile_class == nullptr) return fail(FDTD_ERR_ARG, "material arrays need a tile_class map");
  P.n_slabs = d->n_slabs;
  for (int s = 0; s < d->n_slabs; ++s) P.slabs[s] = slab_k<T, IS_E, A>(d->slabs[s]);
  if (!buf && post_is_fused(d)) {
    P.dyn = graph_step >= 0 ? (const i64*)d->dyn : nullptr;
    for (int n = 0; n < d->n_sources; ++n) {
      const fdtd_source& S = d->sources[n];
      if (S.field != (IS_E ? 0 : 1)) continue;
      if (S.kind == FDTD_SRC_POINTS && S.n == 0) continue;
      int64_t w = q - S.wave_q0;
      if (graph_step < 0 && (w < 0 || w >= S.wave_len))
        return fail(FDTD_ERR_ARG, "source %d: step %lld outside its waveform table [%lld,%lld)", n, (long long)q,
                    (long long)S.wave_q0, (long long)(S.wave_q0 + S.wave_len));
      fdtd::SrcK<A>& K = P.src[P.n_src++];
      K.kind = S.kind;
      K.comp = S.comp;
      K.n = S.n;
      for (int k = 0; k < 6; ++k) K.bb[k] = S.kind == FDTD_SRC_BOX ? S.box[k] : S.bbox[k];
      K.idx = (const i64*)S.idx;
      K.profile = (const A*)S.profile;
      K.amplitude = (A)S.amplitude;
      K.wave = (const A*)S.wave;
      K.w = graph_step >= 0 ? graph_step : w;
    }
    for (int n = 0; n < d->n_detectors; ++n) {
      const fdtd_detector& D = d->detectors[n];
      if (D.n == 0) continue;
      if (graph_step < 0 && (slot < 0 || slot >= D.capacity))
        return fail(FDTD_ERR_ARG, "detector %d: ring slot %lld outside capacity %lld", n, (long long)slot,
                    (long long)D.capacity);
      fdtd::DetK<T>& K = P.det[P.n_det++];
      K.n = D.n;
      for (int k = 0; k < 6; ++k) K.bb[k] = D.bbox[k];
      K.idx = (const i64*)D.idx;
      K.pos = (const int*)D.pos;
      K.ring = (T*)(IS_E ? D.ring_E : D.ring_H);
      K.slot = graph_step >= 0 ? graph_step : slot;
    }
  }

  int chunks = (x_end - x_begin + P.x_chunk - 1) / P.x_chunk;
  dim3 grid((P.z_end - P.z_begin + g.tile_z - 1) / g.tile_z, (P.y_end - P.y_begin + g.tile_y - 1) / g.tile_y, chunks);
  dim3 block(g.lanes_z * g.rows);
  if (grid.y > 65535u || grid.z > 65535u) return fail(FDTD_ERR_UNSUPPORTED, "grid too large for one launch");
  const bool has_post = (P.n_src + P.n_det) > 0;
  const bool has_push = push_y != nullptr && push_z != nullptr;
  if (has_push) {
    P.push_plane = IS_E ? 0 : d->Nx - 1;
    if (P.push_plane < x_begin || P.push_plane >= x_end)
      return fail(FDTD_ERR_ARG, "halo push: the launch does not cover the boundary plane");
    P.push_y = (T*)push_y;
    P.push_z = (T*)push_z;
  }
  // folded sources / detectors: their own material-free instantiation on homogeneous unsharded grids (mid-size
  // grids, where the fold and the graph replay it enables save the launches between the half-steps); otherwise
  // they share the general (MAT) one
  const bool mat = P.cls != nullptr || (has_post && has_push);
#define FDTD_LAUNCH_HALFSTEP(V)                                                                         \
  if (has_post && has_push) {                                                                           \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, true, true, true, A>), grid, block, stream, P);     \
  } else if (has_post && !mat) {                                                                        \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, true, false, false, A>), grid, block, stream, P);   \
  } else if (has_post) {                                                                                \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, true, false, true, A>), grid, block, stream, P);    \
  } else if (has_push && mat) {                                                                         \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, false, true, true, A>), grid, block, stream, P);    \
  } else if (has_push) {                                                                                \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, false, true, false, A>), grid, block, stream, P);   \
  } else if (mat) {                                                                                     \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, false, false, true, A>), grid, block, stream, P);   \
  } else {                                                                                              \
    FDTD_LAUNCH((fdtd::halfstep_kernel<T, V, IS_E, false, false, false, A>), grid, block, stream, P);  \
  }
  switch (g.vec) {
    case 4:
      if constexpr (sizeof(T) == 4) {
        FDTD_LAUNCH_HALFSTEP(4)
      }
      break;
    case 2:
      FDTD_LAUNCH_HALFSTEP(2)
      break;
    default:
      FDTD_LAUNCH_HALFSTEP(1)
  }
#undef FDTD_LAUNCH_HALFSTEP
  return check_launch(IS_E ? "e_halfstep" : "h_halfstep");
}

int blocks_for(i64 n, int threads = 256) {
  i64 b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

// `phases` (FDTD_PHASE_*): which parts of what follows the half-step kernel to run.  With a periodic x boundary
// across slabs (d->x_wrap) the boundary post ops are split at it: BEFORE = those registered before it, AFTER = the
// rest; without one, BEFORE covers them all.
template <typename T, bool IS_E, typename A = T>
int launch_post(const fdtd_desc* d, int64_t q, int64_t slot, void* stream, unsigned phases = FDTD_PHASE_ALL) {
  if (post_is_fused(d)) return FDTD_OK;  // done inside the half-step kernel
  if (d->x_wrap && (phases & FDTD_PHASE_BEFORE) && (phases & FDTD_PHASE_AFTER))
    return fail(FDTD_ERR_ARG, "periodic x boundary across slabs: run the phases around the plane transfer");
  const int split = d->x_wrap ? d->x_wrap - 1 : d->n_post;
  const int first = (phases & FDTD_PHASE_BEFORE) ? 0 : split;
  const int last = (phases & FDTD_PHASE_AFTER) ? d->n_post : ((phases & FDTD_PHASE_BEFORE) ? split : first);
  T* F[3];
  for (int c = 0; c < 3; ++c) F[c] = (T*)(IS_E ? d->E[c] : d->H[c]);
  // 0. objects beyond the second one on a cell, registration order (fdtd/grid.py:285-287); update_H of every
  //    object kind is empty (fdtd/objects.py:131-137, 223-229, 271-277)
  for (int n = 0; IS_E && (phases & FDTD_PHASE_OBJECTS) && n < d->n_deep; ++n) {
    const fdtd_deep_object& O = d->deep[n];
    const i64 cells = (i64)(O.box[1] - O.box[0]) * (O.box[3] - O.box[2]) * (O.box[5] - O.box[4]);
    if (O.box[0] >= O.box[1] || O.box[2] >= O.box[3] || O.box[4] >= O.box[5]) continue;
    FDTD_LAUNCH((fdtd::object_layer_kernel<T, A>), dim3(blocks_for(cells)), dim3(256), stream, F[0], F[1], F[2],
                (const T*)d->H[0], (const T*)d->H[1], (const T*)d->H[2], (const A*)O.inv[0], (const A*)O.inv[1],
                (const A*)O.inv[2], (const A*)O.absorb[0], (const A*)O.absorb[1], (const A*)O.absorb[2], O.mask, O.kind,
                O.box[0], O.box[1], O.box[2], O.box[3], O.box[4], O.box[5], d->Nz, d->plane, d->x_offset,
                (A)d->courant);
    int rc = check_launch("object layer");
    if (rc) return rc;
  }
  // 1. periodic copies and late PML corrections, registration order (fdtd/grid.py:290-291, 316-317)
  for (int n = first; n < last; ++n) {
    if (d->post_kind[n] == FDTD_POST_PERIODIC) {
      int axis = d->post_arg[n];
      int N = axis == 0 ? d->Nx : (axis == 1 ? d->Ny : d->Nz);
      if (N < 2) continue;  // E[0] = E[-1] on a one-cell axis is the identity
      int src = IS_E ? N - 1 : 0, dst = IS_E ? 0 : N - 1;
      i64 cells = (axis == 0 ? (i64)d->Ny * d->Nz : (axis == 1 ? (i64)d->Nx * d->Nz : (i64)d->Nx * d->Ny));
      FDTD_LAUNCH((fdtd::periodic_kernel<T>), dim3(blocks_for(cells * 3)), dim3(256), stream, F[0], F[1], F[2],
                  axis, d->Nx, d->Ny, d->Nz, d->plane, src, dst);
      int rc = check_launch("periodic");
      if (rc) return rc;
    } else {
      const fdtd_slab& S = d->slabs[d->post_arg[n]];
      if (S.psi_count == 0) continue;
      const A* c[3];
      A bg[3];
      for (int k = 0; k < 3; ++k) {
        if (IS_E)
          c[k] = (const A*)(d->inv_eps_grid[k] ? d->inv_eps_grid[k] : d->inv_eps[k]);
        else
          c[k] = (const A*)d->inv_mu[k];
        bg[k] = rounded_product<A>(d->courant, IS_E ? d->bg_inv_eps[k] : d->bg_inv_mu[k]);
      }
      FDTD_LAUNCH((fdtd::pml_add_kernel<T, IS_E, A>), dim3(blocks_for(S.psi_count)), dim3(256), stream,
                  slab_k<T, IS_E, A>(S), F[0], F[1], F[2], c[0], c[1], c[2], bg[0], bg[1], bg[2], (A)d->courant,
                  d->Nx, d->Ny, d->Nz, d->plane);
      int rc = check_launch("pml_add");
      if (rc) return rc;
    }
  }
  // 2. sources, registration order (fdtd/grid.py:294-295, 320-321)
  for (int n = 0; (phases & FDTD_PHASE_SOURCES) && n < d->n_sources; ++n) {
    const fdtd_source& S = d->sources[n];
    if (S.field != (IS_E ? 0 : 1)) continue;
    int64_t w = q - S.wave_q0;
    if (w < 0 || w >= S.wave_len)
      return fail(FDTD_ERR_ARG, "source %d: step %lld outside its waveform table [%lld,%lld)", n, (long long)q,
                  (long long)S.wave_q0, (long long)(S.wave_q0 + S.wave_len));
    if (S.kind == FDTD_SRC_POINTS) {
      if (S.n == 0) continue;
      FDTD_LAUNCH((fdtd::source_points_kernel<T, A>), dim3(blocks_for(S.n)), dim3(256), stream, F[S.comp],
                  (const i64*)S.idx, (const A*)S.profile, S.n, (const A*)S.wave, (i64)w);
    } else if (S.kind == FDTD_SRC_FEEDBACK) {
      if (S.n == 0) continue;
      if (S.record && (slot < 0 || slot >= S.record_capacity))
        return fail(FDTD_ERR_ARG, "source %d: record slot %lld outside capacity", n, (long long)slot);
      i64 cell = (i64)S.box[0] * d->plane + (i64)S.box[2] * d->Nz + S.box[4];
      FDTD_LAUNCH((fdtd::source_feedback_kernel<T, A>), dim3(1), dim3(32), stream, F[2], cell, (const A*)S.wave,
                  (const A*)S.profile, (i64)w, (A)S.impedance, (const T*)S.feedback, (int)(q > 0), (A)S.spacing,
                  (T*)S.record, (i64)slot);
    } else {
      i64 cells = (i64)(S.box[1] - S.box[0]) * (S.box[3] - S.box[2]) * (S.box[5] - S.box[4]);
      if (cells <= 0) continue;
      FDTD_LAUNCH((fdtd::source_box_kernel<T, A>), dim3(blocks_for(cells)), dim3(256), stream, F[S.comp], S.box[0],
                  S.box[1], S.box[2], S.box[3], S.box[4], S.box[5], d->Nz, d->plane, (A)S.amplitude,
                  (const A*)S.wave, (i64)w);
    }
    int rc = check_launch("source");
    if (rc) return rc;
  }
  // 3. detectors (fdtd/grid.py:298-299, 324-325)
  for (int n = 0; (phases & FDTD_PHASE_DETECTORS) && n < d->n_detectors; ++n) {
    const fdtd_detector& D = d->detectors[n];
    if (D.n == 0) continue;
    if (slot < 0 || slot >= D.capacity)
      return fail(FDTD_ERR_ARG, "detector %d: ring slot %lld outside capacity %lld", n, (long long)slot,
                  (long long)D.capacity);
    if (D.kind == FDTD_DET_CURRENT) {
      if (IS_E) continue;  // CurrentDetector.detect_E is empty (fdtd/detectors.py:414-415)
      FDTD_LAUNCH((fdtd::current_kernel<T, A>), dim3(blocks_for(D.n)), dim3(256), stream, (const T*)d->H[0],
                  (const T*)d->H[1], (const i64*)D.idx, (const int*)D.pos, D.n, d->Nx, d->Ny, d->Nz, d->plane,
                  (A)D.spacing, (T*)D.ring_H, (T*)D.last, (i64)slot, (int)(d->x_offset > 0 || d->h_wrap_ghost));
      int rc2 = check_launch("current detector");
      if (rc2) return rc2;
      continue;
    }
    FDTD_LAUNCH((fdtd::detector_kernel<T>), dim3(blocks_for((i64)D.n * 3)), dim3(256), stream, F[0], F[1], F[2],
                (const i64*)D.idx, (const int*)D.pos, D.n, (T*)(IS_E ? D.ring_E : D.ring_H), (i64)slot);
    int rc = check_launch("detector");
    if (rc) return rc;
  }
  return FDTD_OK;
}

}  // namespace

extern "C" {

int32_t fdtd_abi_version(void) { return FDTD_ABI_VERSION; }

int64_t fdtd_sizeof_desc(void) { return (int64_t)sizeof(fdtd_desc); }

int64_t fdtd_sizeof_halo(void) { return (int64_t)sizeof(fdtd_halo); }

const char* fdtd_last_error(void) { return g_err; }

int64_t fdtd_launch_count(void) { return g_launches.load(); }

int fdtd_tile_shape(int32_t dtype, int32_t Ny, int32_t Nz, int32_t* tile_y, int32_t* tile_z) {
  if ((dtype != FDTD_F32 && dtype != FDTD_F64 && dtype != FDTD_F32X) || Ny < 1 || Nz < 1 || !tile_y || !tile_z)
    return fail(FDTD_ERR_ARG, "fdtd_tile_shape: bad argument");
  Geometry g = geometry(dtype, Ny, Nz);
  *tile_y = g.tile_y;
  *tile_z = g.tile_z;
  return FDTD_OK;
}

int fdtd_validate(const fdtd_desc* d) { return validate(d); }

int fdtd_post_is_fused(const fdtd_desc* d) {
  int rc = validate(d);
  if (rc) return rc;
  return post_is_fused(d) ? 1 : 0;
}

int fdtd_e_halfstep(const fdtd_desc* d, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_halfstep<float, true, float>(d, x_begin, x_end, q, slot, stream) : d->dtype == FDTD_F64 ? launch_halfstep<double, true, double>(d, x_begin, x_end, q, slot, stream) : launch_halfstep<float, true, double>(d, x_begin, x_end, q, slot, stream));
}

int fdtd_h_halfstep(const fdtd_desc* d, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_halfstep<float, false, float>(d, x_begin, x_end, q, slot, stream) : d->dtype == FDTD_F64 ? launch_halfstep<double, false, double>(d, x_begin, x_end, q, slot, stream) : launch_halfstep<float, false, double>(d, x_begin, x_end, q, slot, stream));
}

int fdtd_post_E(const fdtd_desc* d, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_post<float, true, float>(d, q, slot, stream) : d->dtype == FDTD_F64 ? launch_post<double, true, double>(d, q, slot, stream) : launch_post<float, true, double>(d, q, slot, stream));
}

int fdtd_post_H(const fdtd_desc* d, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_post<float, false, float>(d, q, slot, stream) : d->dtype == FDTD_F64 ? launch_post<double, false, double>(d, q, slot, stream) : launch_post<float, false, double>(d, q, slot, stream));
}

int fdtd_post_phases(const fdtd_desc* d, int32_t field, uint32_t phases, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if ((field != 0 && field != 1) || (phases & ~(uint32_t)FDTD_PHASE_ALL)) return fail(FDTD_ERR_ARG, "fdtd_post_phases: field / phases");
  if (field == 0)
    return (d->dtype == FDTD_F32 ? launch_post<float, true, float>(d, q, slot, stream, phases) : d->dtype == FDTD_F64 ? launch_post<double, true, double>(d, q, slot, stream, phases) : launch_post<float, true, double>(d, q, slot, stream, phases));
  return (d->dtype == FDTD_F32 ? launch_post<float, false, float>(d, q, slot, stream, phases) : d->dtype == FDTD_F64 ? launch_post<double, false, double>(d, q, slot, stream, phases) : launch_post<float, false, double>(d, q, slot, stream, phases));
}

int fdtd_post_part(const fdtd_desc* d, int32_t field, int32_t part, int64_t q, int64_t slot, void* stream) {
  if (part != 0 && part != 1) return fail(FDTD_ERR_ARG, "fdtd_post_part: part");
  if (d && !d->x_wrap) return fail(FDTD_ERR_ARG, "fdtd_post_part needs d->x_wrap");
  return fdtd_post_phases(d, field, part == 0 ? (FDTD_PHASE_OBJECTS | FDTD_PHASE_BEFORE)
                                              : (FDTD_PHASE_AFTER | FDTD_PHASE_SOURCES | FDTD_PHASE_DETECTORS),
                          q, slot, stream);
}

static int update_E_nocheck(const fdtd_desc* d, int64_t q, int64_t slot, void* stream, int64_t graph_step = -1) {
  int rc = (d->dtype == FDTD_F32 ? launch_halfstep<float, true, float>(d, 0, d->Nx, q, slot, stream, graph_step) : d->dtype == FDTD_F64 ? launch_halfstep<double, true, double>(d, 0, d->Nx, q, slot, stream, graph_step) : launch_halfstep<float, true, double>(d, 0, d->Nx, q, slot, stream, graph_step));
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_post<float, true, float>(d, q, slot, stream) : d->dtype == FDTD_F64 ? launch_post<double, true, double>(d, q, slot, stream) : launch_post<float, true, double>(d, q, slot, stream));
}

static int update_H_nocheck(const fdtd_desc* d, int64_t q, int64_t slot, void* stream, int64_t graph_step = -1) {
  int rc = (d->dtype == FDTD_F32 ? launch_halfstep<float, false, float>(d, 0, d->Nx, q, slot, stream, graph_step) : d->dtype == FDTD_F64 ? launch_halfstep<double, false, double>(d, 0, d->Nx, q, slot, stream, graph_step) : launch_halfstep<float, false, double>(d, 0, d->Nx, q, slot, stream, graph_step));
  if (rc) return rc;
  return (d->dtype == FDTD_F32 ? launch_post<float, false, float>(d, q, slot, stream) : d->dtype == FDTD_F64 ? launch_post<double, false, double>(d, q, slot, stream) : launch_post<float, false, double>(d, q, slot, stream));
}

int fdtd_update_E(const fdtd_desc* d, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return update_E_nocheck(d, q, slot, stream);
}

int fdtd_update_H(const fdtd_desc* d, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  return update_H_nocheck(d, q, slot, stream);
}

// ---- direct peer-to-peer halo exchange (x-sharded grids, one process per GPU) ----------------------------
int fdtd_halfstep_push(const fdtd_desc* d, int32_t field, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot,
                       void* peer_ghost_y, void* peer_ghost_z, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!peer_ghost_y || !peer_ghost_z) return fail(FDTD_ERR_ARG, "null peer ghost pointer");
  if (field != 0 && field != 1) return fail(FDTD_ERR_ARG, "field must be 0 (E) or 1 (H)");
  if (field == 0)
    return (d->dtype == FDTD_F32 ? launch_halfstep<float, true, float>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z) : d->dtype == FDTD_F64 ? launch_halfstep<double, true, double>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z) : launch_halfstep<float, true, double>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z));
  return (d->dtype == FDTD_F32 ? launch_halfstep<float, false, float>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z) : d->dtype == FDTD_F64 ? launch_halfstep<double, false, double>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z) : launch_halfstep<float, false, double>(d, x_begin, x_end, q, slot, stream, -1, peer_ghost_y, peer_ghost_z));
}

int fdtd_halo_push(const fdtd_desc* d, int32_t field, void* peer_ghost_y, void* peer_ghost_z, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!peer_ghost_y || !peer_ghost_z) return fail(FDTD_ERR_ARG, "null peer ghost pointer");
  if (field != 0 && field != 1) return fail(FDTD_ERR_ARG, "field must be 0 (E) or 1 (H)");
  const int64_t plane_off = field == 0 ? 0 : (int64_t)(d->Nx - 1) * d->plane;
  void* const* F = field == 0 ? d->E : d->H;
  if (d->dtype != FDTD_F64) {
    FDTD_LAUNCH((fdtd::halo_push_kernel<float>), dim3(blocks_for(d->plane)), dim3(256), stream,
                (const float*)F[1] + plane_off, (const float*)F[2] + plane_off, (float*)peer_ghost_y,
                (float*)peer_ghost_z, (i64)d->plane);
  } else {
    FDTD_LAUNCH((fdtd::halo_push_kernel<double>), dim3(blocks_for(d->plane)), dim3(256), stream,
                (const double*)F[1] + plane_off, (const double*)F[2] + plane_off, (double*)peer_ghost_y,
                (double*)peer_ghost_z, (i64)d->plane);
  }
  return check_launch("halo_push");
}

int fdtd_dft_accumulate(int32_t dtype, const void* ring, int64_t n_steps, int64_t n_values, const double* twiddle,
                        int32_t n_freqs, double* acc, void* stream) {
  if (dtype != FDTD_F32 && dtype != FDTD_F64 && dtype != FDTD_F32X) return fail(FDTD_ERR_ARG, "bad dtype %d", dtype);
  if (n_steps < 0 || n_values < 0 || n_freqs < 0) return fail(FDTD_ERR_ARG, "fdtd_dft_accumulate: negative size");
  if (n_steps == 0 || n_values == 0 || n_freqs == 0) return FDTD_OK;
  if (!ring || !twiddle || !acc) return fail(FDTD_ERR_ARG, "fdtd_dft_accumulate: null pointer");
  const dim3 grid(blocks_for(n_values * n_freqs)), block(256);
  if (dtype != FDTD_F64) {
    FDTD_LAUNCH((fdtd::dft_accumulate_kernel<float>), grid, block, stream, (const float*)ring, (i64)n_steps,
                (i64)n_values, twiddle, (int)n_freqs, acc);
  } else {
    FDTD_LAUNCH((fdtd::dft_accumulate_kernel<double>), grid, block, stream, (const double*)ring, (i64)n_steps,
                (i64)n_values, twiddle, (int)n_freqs, acc);
  }
  return check_launch("dft_accumulate");
}

#ifdef FDTD_EMU
int fdtd_halo_signal(int64_t*, int64_t, void*) { return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA"); }
int fdtd_halo_wait(const int64_t*, int64_t, int32_t*, int64_t, void*) { return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA"); }
int fdtd_ipc_export(const void*, void*, int64_t*) { return fail(FDTD_ERR_UNSUPPORTED, "CUDA IPC needs CUDA"); }
int fdtd_ipc_import(const void*, int64_t, void**) { return fail(FDTD_ERR_UNSUPPORTED, "CUDA IPC needs CUDA"); }
#else
int fdtd_halo_signal(int64_t* peer_flag, int64_t value, void* stream) {
  if (!peer_flag) return fail(FDTD_ERR_ARG, "null flag");
  FDTD_LAUNCH((fdtd::halo_signal_kernel), dim3(1), dim3(32), stream, (i64*)peer_flag, (i64)value);
  return check_launch("halo_signal");
}

int fdtd_halo_wait(const int64_t* flag, int64_t value, int32_t* error, int64_t timeout_ns, void* stream) {
  if (!flag || !error) return fail(FDTD_ERR_ARG, "null flag");
  FDTD_LAUNCH((fdtd::halo_wait_kernel), dim3(1), dim3(32), stream, (const i64*)flag, (i64)value, (int*)error,
              (i64)timeout_ns);
  return check_launch("halo_wait");
}

int fdtd_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
  if (!dev_ptr || !handle64 || !offset) return fail(FDTD_ERR_ARG, "fdtd_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  // the handle names the whole cudaMalloc allocation: find its base through the driver API (resolved at run
  // time so that the library loads without libcuda in CPU-only containers)
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return fail(FDTD_ERR_CUDA, "cuMemGetAddressRange not available");
    get_range = (range_fn)fn;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
    return fail(FDTD_ERR_CUDA, "cuMemGetAddressRange failed");
  cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, (void*)(uintptr_t)base);
  if (e != cudaSuccess) return fail(FDTD_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  *offset = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return FDTD_OK;
}

int fdtd_ipc_import(const void* handle64, int64_t offset, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return fail(FDTD_ERR_ARG, "fdtd_ipc_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(FDTD_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  *dev_ptr = (char*)base + offset;
  return FDTD_OK;
}
#endif

// ---- temporally fused E+H steps (yee_fused_eh.cuh) --------------------------------------------------------
extern "C++" {
namespace {
// automatic mode (fuse_eh = 2): the fused kernel wins where its 7 x 31-vector tiles quantise the y-z plane well and
// the march is long enough -- measured on the B200 (profiles/r2_s14/fused_sizes.log): float32 512^3 -5 % (4.13 z tiles),
// 576^3 +12 %, 640^3 +7 %, 768^3 +13 %, 1024^3 +25 %; float64 256^3 -18 %, 384^3 +6 %, 512^3 +19 %; slabs of 1024^2
// planes: 64 planes +2 %, 96 +3 %, 128 +10 %, 256 +17 %
#ifdef FDTD_EMU
#define FDTD_FUSE_EH_MIN_PLANE_BYTES 0
#define FDTD_FUSE_EH_MIN_PLANES 2
#define FDTD_FUSE_EH_MIN_Z_FILL 0.0
#else
#define FDTD_FUSE_EH_MIN_PLANE_BYTES (1100LL << 10)
#define FDTD_FUSE_EH_MIN_PLANES 64
#define FDTD_FUSE_EH_MIN_Z_FILL 0.85     // Nz / (z tiles x tile length): the last z tile of a row is mostly empty below
#endif

// fuse_eh = 1: wherever it is legal; fuse_eh = 2: only where it is also faster (large grids: what counts is the
// tile quantisation of the y-z plane, so an x-slab of a large grid qualifies like the grid itself)
bool fuse_eh_eligible(const fdtd_desc* d, bool sharded = false) {
  if (!d->fuse_eh || d->dtype == FDTD_F32X) return false;
  for (int c = 0; c < 3; ++c)
    if (!d->E2[c] || !d->H2[c] || d->inv_eps[c] || d->inv_mu[c] || d->absorb[c]) return false;
  if ((d->Nx != d->Nx_global) != sharded || d->n_post != 0 || d->n_deep != 0 || d->x_wrap) return false;
  const int vec = d->dtype == FDTD_F32 ? 4 : 2;
  if (d->Nz % vec) return false;
  const int tile_z = fdtd::FUSED_L * vec;
  if (d->fuse_eh == 2 && ((int64_t)d->Ny * d->Nz * (d->dtype == FDTD_F32 ? 4 : 8) < FDTD_FUSE_EH_MIN_PLANE_BYTES ||
                          d->Nx < FDTD_FUSE_EH_MIN_PLANES ||
                          (double)d->Nz < FDTD_FUSE_EH_MIN_Z_FILL * (double)((d->Nz + tile_z - 1) / tile_z * tile_z)))
    return false;
  if (sharded && d->Nx < 4) return false;
  int nsrc = 0;
  for (int n = 0; n < d->n_sources; ++n) {
    if (d->sources[n].kind != FDTD_SRC_POINTS || d->sources[n].field != 0) return false;
    ++nsrc;
  }
  if (nsrc > FDTD_FUSED_MAX) return false;
  for (int n = 0; n < d->n_detectors; ++n)
    if (d->detectors[n].kind != FDTD_DET_FIELD) return false;
  for (int s = 0; s < d->n_slabs; ++s) {
    // the kernel applies every CPML correction itself (slabs registered after a periodic boundary are post ops, and
    // periodic boundaries are excluded above anyway) and needs the second psi_E buffer
    if (!d->slabs[s].fused || (d->slabs[s].psi_count > 0 && !d->psi_E2[s])) return false;
    if (d->slabs[s].thickness > FDTD_FUSED_TAB_T) return false;   // (its coefficient tables live in shared memory)
  }
#ifndef FDTD_EMU
  return d->Nx >= 8 && d->Ny >= 8 && d->Nz >= 32 * vec;
#else
  return d->Nx >= 2;     // (CPU tests: small grids with partial tiles exercise every branch)
#endif
}

// tensor maps of the three components of one field buffer for the fused kernel's staging: a 3-D tensor
// [Nx + 2][Ny][Nz] per component (the ghost x-planes belong to it), box = (bz x by x 1 plane)
template <typename T>
int fused_tma_maps(const fdtd_desc* d, void* const* F, int bz, int by, fdtd::TmaMap<T>* out) {
#ifdef FDTD_EMU
  for (int c = 0; c < 3; ++c) {
    out[c].base = (const T*)F[c] - d->plane;
    out[c].n0 = d->Nz;
    out[c].n1 = d->Ny;
    out[c].n2 = d->Nx + 2;
  }
  return FDTD_OK;
#else
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return fail(FDTD_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (encode_fn)fn;
  }
  // a handful of buffers (two per field): keep the encoded maps
  struct Entry {
    void* base;
    int Nx, Ny, Nz, bz, by;
    CUtensorMap m;
  };
  static thread_local Entry cache[16];
  static thread_local int next = 0;
  for (int c = 0; c < 3; ++c) {
    void* base = (void*)((T*)F[c] - d->plane);
    const Entry* hit = nullptr;
    for (const Entry& e : cache)
      if (e.base == base && e.Nx == d->Nx && e.Ny == d->Ny && e.Nz == d->Nz && e.bz == bz && e.by == by) hit = &e;
    if (!hit) {
      Entry& e = cache[next];
      next = (next + 1) % 16;
      const cuuint64_t dims[3] = {(cuuint64_t)d->Nz, (cuuint64_t)d->Ny, (cuuint64_t)d->Nx + 2};
      const cuuint64_t strides[2] = {(cuuint64_t)d->Nz * sizeof(T), (cuuint64_t)d->plane * sizeof(T)};
      const cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)by, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      // (development knob: FDTD_B200_TMA_L2PROMO = 0 none / 1 64 B / 2 128 B / 3 256 B)
      static const int promo = [] { const char* v = getenv("FDTD_B200_TMA_L2PROMO"); return v ? atoi(v) : 2; }();
      const CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                        : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                        : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                     : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
      CUresult r = encode(&e.m, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base,
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        e.base = nullptr;
        return fail(FDTD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
      }
      e.base = base; e.Nx = d->Nx; e.Ny = d->Ny; e.Nz = d->Nz; e.bz = bz; e.by = by;
      hit = &e;
    }
    out[c].m = hit->m;
  }
  return FDTD_OK;
#endif
}

// what an x-sharded slab adds to a fused step
struct FusedShard {
  void* push_y;        // the left neighbour's ghost planes (in ITS output buffer) for E_new[plane 0], or null
  void* push_z;
  int skip_last_h;     // the H update of the last local plane is a separate launch (it waits for the right neighbour)
  int part;            // 0: every x-chunk; 1: only the first and the last one (the planes the neighbours wait for);
                       // 2: only the ones in between
  int* n_chunks;       // out (may be null): how many x-chunks the slab is cut into
};

template <typename T>
int fused_detectors(const fdtd_desc* d, void* const* Eout, void* const* Hout, int64_t slot, void* stream) {
  for (int n = 0; n < d->n_detectors; ++n) {
    const fdtd_detector& D = d->detectors[n];
    if (D.n == 0) continue;
    if (slot < 0 || slot >= D.capacity) return fail(FDTD_ERR_ARG, "detector %d: ring slot outside capacity", n);
    FDTD_LAUNCH((fdtd::detector_kernel<T>), dim3(blocks_for((i64)D.n * 3)), dim3(256), stream, (const T*)Eout[0],
                (const T*)Eout[1], (const T*)Eout[2], (const i64*)D.idx, (const int*)D.pos, D.n, (T*)D.ring_E,
                (i64)slot);
    int rc = check_launch("detector");
    if (rc) return rc;
    FDTD_LAUNCH((fdtd::detector_kernel<T>), dim3(blocks_for((i64)D.n * 3)), dim3(256), stream, (const T*)Hout[0],
                (const T*)Hout[1], (const T*)Hout[2], (const i64*)D.idx, (const int*)D.pos, D.n, (T*)D.ring_H,
                (i64)slot);
    rc = check_launch("detector");
    if (rc) return rc;
  }
  return FDTD_OK;
}

// one full step reading (Ein, Hin) and writing (Eout, Hout); parity 0: psi_E -> psi_E2, 1: back
template <typename T>
int fused_eh_step(const fdtd_desc* d, void* const* Ein, void* const* Eout, void* const* Hin, void* const* Hout,
                  int64_t q, int64_t slot, void* stream, int parity, const FusedShard* shard = nullptr) {
  const int Nx = d->Nx, Ny = d->Ny, Nz = d->Nz;
  int rc;
  fdtd::FusedParams<T> P;
  memset(&P, 0, sizeof(P));
  for (int n = 0; n < d->n_sources; ++n) {
    const fdtd_source& S = d->sources[n];
    int64_t w = q - S.wave_q0;
    if (w < 0 || w >= S.wave_len)
      return fail(FDTD_ERR_ARG, "source %d: step %lld outside its waveform table", n, (long long)q);
    if (S.n == 0) continue;
    fdtd::SrcK<T>& K = P.src[P.n_src++];
    K.kind = S.kind;
    K.comp = S.comp;
    K.n = S.n;
    for (int k = 0; k < 6; ++k) K.bb[k] = S.bbox[k];
    K.idx = (const i64*)S.idx;
    K.profile = (const T*)S.profile;
    K.wave = (const T*)S.wave;
    K.w = w;
  }
  constexpr int VEC = sizeof(T) == 4 ? 4 : 2;
  P.Nx = Nx;
  P.Ny = Ny;
  P.Nz = Nz;
  P.plane = d->plane;
  P.x_offset = d->x_offset;
  P.Nx_global = d->Nx_global;
  if (shard) {
    P.push_y = (T*)shard->push_y;
    P.push_z = (T*)shard->push_z;
    P.skip_last_h = shard->skip_last_h;
  }
  P.x0 = 0; P.x1 = Nx; P.y0 = 0; P.y1 = Ny; P.z0 = 0; P.z1 = Nz;
  P.psi_stage = 1;
  // The march along x is cut into chunks of ~40 planes (1024^3 f32: 9.99 / 9.95 / 9.90 / 9.92 / 9.90 / 10.01 ms per step
  // at 24 / 32 / 36 / 40 / 44 / 49 planes, profiles/r2_s19/): every chunk costs a pipeline fill and one extra E plane,
  // and the blocks that run last finish with the GPU half empty -- so ONE chunk is only about half as long as the
  // others and its blocks are launched last (128 planes: 37+37+37+17 runs 3 % faster than 4 x 32).  On an x-sharded slab
  // the first and the last chunk run first (fused_sharded_step): the short chunk is the second to last there, the end
  // of what the caller's stream runs.  grid._x_chunk / desc.x_chunk > 0: chunks of exactly that length.
  const bool split = shard && shard->part != 0;
  int n_chunks = 0;
  {
    const int cap = FDTD_FUSED_MAX_CHUNKS;
    int len, c, short_at = -1, short_len = 0;
    if (d->x_chunk > 0) {
      len = d->x_chunk;
      if ((Nx + len - 1) / len > cap) len = (Nx + cap - 1) / cap;
      c = (Nx + len - 1) / len;
    } else {
#ifdef FDTD_EMU
      const double target = 5.0;                     // (CPU tests: small grids get several chunks, too)
#else
      const double target = 40.0;
#endif
      c = (int)((double)Nx / target + 0.999);        // chunks of ~40 planes, one of them half as long
      if (c > cap) c = cap;
      if (c < 1) c = 1;
      len = c > 1 ? (int)((double)Nx / (c - 0.5) + 0.999) : Nx;
      while (c > 1 && (c - 1) * len >= Nx) --len;    // (the short chunk must not be empty)
      short_len = Nx - (c - 1) * len;
      short_at = (split && c >= 3) ? c - 2 : c - 1;
    }
    int x = 0;
    for (int k = 0; k < c; ++k) {
      P.xstart[k] = x;
      x += (k == short_at) ? short_len : len;
      if (x > Nx) x = Nx;
    }
    P.xstart[c] = Nx;
    n_chunks = c;
  }
  for (int c = 0; c < 3; ++c) {
    P.Ein[c] = (const T*)Ein[c];
    P.Eout[c] = (T*)Eout[c];
    P.Hin[c] = (const T*)Hin[c];
    P.Hout[c] = (T*)Hout[c];
    P.ce[c] = rounded_product<T>(d->courant, d->bg_inv_eps[c]);
    P.ch[c] = rounded_product<T>(d->courant, d->bg_inv_mu[c]);
  }
  // every slab inside the kernel, registration order: psi_E read from one buffer and written to the other
  for (int s = 0; s < d->n_slabs; ++s) {
    const fdtd_slab& S = d->slabs[s];
    if (S.psi_count == 0) continue;
    typename fdtd::FusedParams<T>::Slab& K = P.sl[P.n_sl++];
    K.axis = S.axis;
    K.lo = S.axis == 0 ? S.lo - d->x_offset : S.lo;
    K.xs = S.x0;
    K.xe = S.x1;
    K.t = S.thickness;
    K.lo_al = z_slab_lo(S);
    K.tp = z_slab_row(S);
    K.count = S.psi_count;
    K.psiE_in = (const T*)(parity == 0 ? S.psi_E : d->psi_E2[s]);
    K.psiE_out = (T*)(parity == 0 ? d->psi_E2[s] : S.psi_E);
    K.psiH = (T*)S.psi_H;
    K.bE = (const T*)S.bE;
    K.cE = (const T*)S.cE;
    K.bH = (const T*)S.bH;
    K.cH = (const T*)S.cH;
    // (bulk copies of psi rows: 16-byte aligned sources; rows and array halves are multiples of four words)
    if (S.axis == 2 && !(aligned(K.psiE_in, 16) && aligned(K.psiH, 16))) P.psi_stage = 0;
  }
  using Lay = fdtd::FusedPipeLayout<T, VEC>;
  const int chunks = n_chunks;
  int launch_chunks = chunks;
  P.chunk0 = 0;
  P.chunk_step = 1;
  if (shard && shard->n_chunks) *shard->n_chunks = chunks;
  if (shard && shard->part == 1) {
    P.chunk_step = chunks > 1 ? chunks - 1 : 1;
    launch_chunks = chunks > 1 ? 2 : 1;
  } else if (shard && shard->part == 2) {
    P.chunk0 = 1;
    launch_chunks = chunks - 2;
    if (launch_chunks <= 0) return FDTD_OK;
  }
  dim3 grid((Nz + fdtd::FUSED_L * VEC - 1) / (fdtd::FUSED_L * VEC), (Ny + fdtd::FUSED_R - 1) / fdtd::FUSED_R,
            (unsigned)launch_chunks);
  dim3 block((fdtd::FUSED_R + 1) * (fdtd::FUSED_L + 1));
  if (grid.y > 65535u || grid.z > 65535u) return fail(FDTD_ERR_UNSUPPORTED, "grid too large for one launch");
  fdtd::FusedTmaMaps<T> M;
  memset(&M, 0, sizeof(M));
  // (a TMA row is a multiple of 16 bytes: Nz % VEC == 0, checked by the eligibility test)
  if ((rc = fused_tma_maps<T>(d, Hin, Lay::HV * VEC, Lay::R + 2, M.h)) != 0) return rc;
  if ((rc = fused_tma_maps<T>(d, Ein, Lay::EV * VEC, Lay::R + 1, M.e)) != 0) return rc;
#ifndef FDTD_EMU
  static bool configured = false;      // (per instantiation: one kernel function each)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fdtd::fused_eh_pipe_kernel<T, VEC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::BYTES);
    if (e != cudaSuccess) return fail(FDTD_ERR_CUDA, "cudaFuncSetAttribute(smem %zu): %s", Lay::BYTES, cudaGetErrorString(e));
    // two blocks per SM only fit with the shared-memory carve-out at its maximum
    cudaFuncSetAttribute(fdtd::fused_eh_pipe_kernel<T, VEC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
#endif
  FDTD_LAUNCH_SMEM((fdtd::fused_eh_pipe_kernel<T, VEC>), grid, block, Lay::BYTES, stream, P, M);
  rc = check_launch("fused_eh");
  if (rc) return rc;
  // detectors on the new fields (a sharded step samples after its last H plane)
  if (!shard) return fused_detectors<T>(d, Eout, Hout, slot, stream);
  return FDTD_OK;
}

// an unsharded fused step; FDTD_B200_FUSE_SPLIT_TEST=1 (tests): as two launches, the first and last x-chunk and then
// the ones in between -- the chunk subsets an x-sharded slab runs on two streams
template <typename T>
int fused_eh_step_unsharded(const fdtd_desc* d, void* const* Ein, void* const* Eout, void* const* Hin, void* const* Hout,
                            int64_t q, int64_t slot, void* stream, int parity) {
  const char* v = getenv("FDTD_B200_FUSE_SPLIT_TEST");
  if (!v || atoi(v) == 0) return fused_eh_step<T>(d, Ein, Eout, Hin, Hout, q, slot, stream, parity);
  FusedShard sh{nullptr, nullptr, 0, 1, nullptr};
  int rc = fused_eh_step<T>(d, Ein, Eout, Hin, Hout, q, slot, stream, parity, &sh);
  if (rc) return rc;
  sh.part = 2;
  if ((rc = fused_eh_step<T>(d, Ein, Eout, Hin, Hout, q, slot, stream, parity, &sh)) != 0) return rc;
  return fused_detectors<T>(d, Eout, Hout, slot, stream);
}
}  // namespace
}  // extern "C++"

// ---- a whole half-step / run of an x-sharded slab on the C side (no host work between the steps) -----------
#ifdef FDTD_EMU
int fdtd_sharded_halfstep(const fdtd_desc*, fdtd_halo*, int32_t, int64_t, int64_t, void*) {
  return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA");
}
int fdtd_run_sharded(const fdtd_desc*, fdtd_halo*, int64_t, int64_t, int64_t, void*) {
  return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA");
}
int fdtd_fuse_eh_sharded_active(const fdtd_desc*, const fdtd_halo*) { return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA"); }
int fdtd_halo_refresh(const fdtd_desc*, fdtd_halo*, void*) { return fail(FDTD_ERR_UNSUPPORTED, "peer-to-peer halo needs CUDA"); }
#else
extern "C++" {
namespace {
// two events order the caller's stream and the side stream; re-recorded every half-step (a wait refers to the
// most recent record at the time it is enqueued)
struct StreamJoin {
  cudaEvent_t to_side = nullptr, to_main = nullptr;
  int ensure() {
    if (!to_side && cudaEventCreateWithFlags(&to_side, cudaEventDisableTiming) != cudaSuccess)
      return fail(FDTD_ERR_CUDA, "cudaEventCreate failed");
    if (!to_main && cudaEventCreateWithFlags(&to_main, cudaEventDisableTiming) != cudaSuccess)
      return fail(FDTD_ERR_CUDA, "cudaEventCreate failed");
    return FDTD_OK;
  }
};
thread_local StreamJoin g_join;

int stream_after(cudaStream_t waiter, cudaStream_t signaller, cudaEvent_t ev) {
  if (cudaEventRecord(ev, signaller) != cudaSuccess || cudaStreamWaitEvent(waiter, ev, 0) != cudaSuccess)
    return fail(FDTD_ERR_CUDA, "stream join: %s", cudaGetErrorString(cudaGetLastError()));
  return FDTD_OK;
}

int check_halo(const fdtd_desc* d, const fdtd_halo* h) {
  if (!h) return fail(FDTD_ERR_ARG, "null halo");
  if (d->Nx == d->Nx_global) return fail(FDTD_ERR_ARG, "not an x-sharded slab");
  if (!h->flags || !h->error || !h->side_stream) return fail(FDTD_ERR_ARG, "halo: null flags / error / side stream");
  if (h->has_left && (!h->left_ghost_y || !h->left_ghost_z || !h->left_flag))
    return fail(FDTD_ERR_ARG, "halo: left neighbour pointers");
  if (h->has_right && (!h->right_ghost_y || !h->right_ghost_z || !h->right_flag))
    return fail(FDTD_ERR_ARG, "halo: right neighbour pointers");
  return FDTD_OK;
}

template <typename T, typename A>
int sharded_halfstep(const fdtd_desc* d, fdtd_halo* h, int field, int64_t q, int64_t slot, void* stream) {
  const int n = d->Nx;
  const bool is_e = field == 0;
  cudaStream_t main = (cudaStream_t)stream, side = (cudaStream_t)h->side_stream;
  // E: plane 0 needs the left neighbour's H ghost and goes to the left neighbour; H: the last plane needs the
  // right neighbour's E ghost and goes to the right neighbour
  const bool has_nb = is_e ? h->has_left != 0 : h->has_right != 0;
  const int b0 = is_e ? 1 : 0, b1 = is_e ? n : n - 1;
  const int e0 = is_e ? 0 : (n - 1 > 0 ? n - 1 : 0), e1 = is_e ? (n < 1 ? n : 1) : n;
  void* gy = is_e ? h->left_ghost_y : h->right_ghost_y;
  void* gz = is_e ? h->left_ghost_z : h->right_ghost_z;
  int64_t* peer_flag = is_e ? h->left_flag : h->right_flag;
  int rc = stream_after(side, main, g_join.to_side);      // everything enqueued so far (user writes included)
  if (rc) return rc;
  rc = is_e ? launch_halfstep<T, true, A>(d, b0, b1, q, slot, main)
            : launch_halfstep<T, false, A>(d, b0, b1, q, slot, main);
  if (rc) return rc;
  if (has_nb) {
    // the ghost this plane needs was pushed by the neighbour after its last half-step of the OTHER field
    FDTD_LAUNCH((fdtd::halo_wait_kernel), dim3(1), dim3(32), side, (const i64*)h->flags + (1 - field),
                (i64)h->count[1 - field], (int*)h->error, (i64)h->timeout_ns);
    rc = check_launch("halo_wait");
    if (rc) return rc;
  }
  const bool fused = has_nb && h->push_fused[field];
  rc = is_e ? launch_halfstep<T, true, A>(d, e0, e1, q, slot, side, -1, fused ? gy : nullptr, fused ? gz : nullptr)
            : launch_halfstep<T, false, A>(d, e0, e1, q, slot, side, -1, fused ? gy : nullptr, fused ? gz : nullptr);
  if (rc) return rc;
  if (fused) {
    FDTD_LAUNCH((fdtd::halo_signal_kernel), dim3(1), dim3(32), side, (i64*)peer_flag, (i64)(h->count[field] + 1));
    rc = check_launch("halo_signal");
    if (rc) return rc;
  }
  rc = stream_after(main, side, g_join.to_main);
  if (rc) return rc;
  rc = is_e ? launch_post<T, true, A>(d, q, slot, main) : launch_post<T, false, A>(d, q, slot, main);
  if (rc) return rc;
  if (has_nb && !fused) {
    rc = stream_after(side, main, g_join.to_side);
    if (rc) return rc;
    const int64_t plane_off = is_e ? 0 : (int64_t)(n - 1) * d->plane;
    void* const* F = is_e ? d->E : d->H;
    FDTD_LAUNCH((fdtd::halo_push_kernel<T>), dim3(blocks_for(d->plane)), dim3(256), side, (const T*)F[1] + plane_off,
                (const T*)F[2] + plane_off, (T*)gy, (T*)gz, (i64)d->plane);
    rc = check_launch("halo_push");
    if (rc) return rc;
    FDTD_LAUNCH((fdtd::halo_signal_kernel), dim3(1), dim3(32), side, (i64*)peer_flag, (i64)(h->count[field] + 1));
    rc = check_launch("halo_signal");
    if (rc) return rc;
  }
  h->count[field] += 1;
  return FDTD_OK;
}
}  // namespace
}  // extern "C++"

extern "C++" {
namespace {
bool fuse_eh_sharded(const fdtd_desc* d, const fdtd_halo* h) {
  if (!fuse_eh_eligible(d, true)) return false;
  if (h->has_left && (!h->left_ghost_y2 || !h->left_ghost_z2)) return false;
  if (h->has_right && (!h->right_ghost_y2 || !h->right_ghost_z2)) return false;
  return true;
}

// One temporally fused step of an x-sharded slab, reading buffer pair (Ein, Hin) and writing (Eout, Hout).
//   1. wait for the left neighbour's last H plane of the PREVIOUS step (the ghost of Hin);
//   2. the fused kernel: E_new everywhere -- plane 0 also stored into the left neighbour's ghost of ITS Eout -- and
//      H_new on all planes but the last one; then the flag that publishes plane 0;
//   3. wait for the right neighbour's E_new[0] in the ghost of Eout, then the H update of the last plane with the
//      ordinary half-step kernel (Hin -> Hout, curls from Eout), stored into the right neighbour's ghost of ITS Hout
//      as well, and the flag that publishes it;
//   4. detectors.
// Flags and counts are the ones of the two-half-step protocol (one E push and one H push per step), so fused and
// ordinary steps can follow each other.
template <typename T>
int fused_sharded_step(const fdtd_desc* d, fdtd_halo* h, int parity, int64_t q, int64_t slot, void* stream) {
  void* const* Ein = parity == 0 ? d->E : d->E2;
  void* const* Eout = parity == 0 ? d->E2 : d->E;
  void* const* Hin = parity == 0 ? d->H : d->H2;
  void* const* Hout = parity == 0 ? d->H2 : d->H;
  cudaStream_t main = (cudaStream_t)stream, side = (cudaStream_t)h->side_stream;
  int rc;
  // The x-chunks the neighbours wait for -- the first one (E_new[plane 0] goes left) and the last one (its E_new feeds
  // the H update of the last plane, which goes right) -- run on the (high-priority) side stream together with the
  // flags and the last H plane; the chunks in between run on the caller's stream meanwhile and hide that serial chain.
  // FDTD_B200_FUSE_SPLIT=0: everything in one launch on the caller's stream.
  static const bool want_split = [] { const char* v = getenv("FDTD_B200_FUSE_SPLIT"); return !v || atoi(v) != 0; }();
  int n_chunks = 0;
  FusedShard sh;
  sh.push_y = h->has_left ? (parity == 0 ? h->left_ghost_y2 : h->left_ghost_y) : nullptr;
  sh.push_z = h->has_left ? (parity == 0 ? h->left_ghost_z2 : h->left_ghost_z) : nullptr;
  sh.skip_last_h = h->has_right;
  sh.n_chunks = &n_chunks;
  sh.part = want_split ? 1 : 0;
  cudaStream_t edge = want_split ? side : main;
  if (want_split) {
    // both streams have seen everything of the previous step
    if ((rc = stream_after(side, main, g_join.to_side)) != 0) return rc;
    if ((rc = stream_after(main, side, g_join.to_main)) != 0) return rc;
  }
  if (h->has_left) {
    FDTD_LAUNCH((fdtd::halo_wait_kernel), dim3(1), dim3(32), edge, (const i64*)h->flags + 1, (i64)h->count[1],
                (int*)h->error, (i64)h->timeout_ns);
    if ((rc = check_launch("halo_wait")) != 0) return rc;
  }
  if ((rc = fused_eh_step<T>(d, Ein, Eout, Hin, Hout, q, slot, edge, parity, &sh)) != 0) return rc;
  if (want_split) {
    sh.part = 2;
    sh.n_chunks = nullptr;
    if ((rc = fused_eh_step<T>(d, Ein, Eout, Hin, Hout, q, slot, main, parity, &sh)) != 0) return rc;
  }
  if (h->has_left) {
    FDTD_LAUNCH((fdtd::halo_signal_kernel), dim3(1), dim3(32), edge, (i64*)h->left_flag, (i64)(h->count[0] + 1));
    if ((rc = check_launch("halo_signal")) != 0) return rc;
  }
  h->count[0] += 1;
  if (h->has_right) {
    FDTD_LAUNCH((fdtd::halo_wait_kernel), dim3(1), dim3(32), edge, (const i64*)h->flags, (i64)h->count[0],
                (int*)h->error, (i64)h->timeout_ns);
    if ((rc = check_launch("halo_wait")) != 0) return rc;
    Buffers buf{Hin, Hout, Eout};
    rc = launch_halfstep_run<T, false, T>(d, d->Nx - 1, d->Nx, q, slot, edge, -1,
                                          parity == 0 ? h->right_ghost_y2 : h->right_ghost_y,
                                          parity == 0 ? h->right_ghost_z2 : h->right_ghost_z, true, &buf);
    if (rc) return rc;
    FDTD_LAUNCH((fdtd::halo_signal_kernel), dim3(1), dim3(32), edge, (i64*)h->right_flag, (i64)(h->count[1] + 1));
    if ((rc = check_launch("halo_signal")) != 0) return rc;
  }
  h->count[1] += 1;
  if (want_split && (rc = stream_after(main, side, g_join.to_main)) != 0) return rc;
  return fused_detectors<T>(d, Eout, Hout, slot, stream);
}
}  // namespace
}  // extern "C++"

int fdtd_fuse_eh_sharded_active(const fdtd_desc* d, const fdtd_halo* h) {
  int rc = validate(d);
  if (rc) return rc;
  if ((rc = check_halo(d, h)) != 0) return rc;
  return fuse_eh_sharded(d, h) ? 1 : 0;
}

int fdtd_sharded_halfstep(const fdtd_desc* d, fdtd_halo* h, int32_t field, int64_t q, int64_t slot, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if ((rc = check_halo(d, h)) != 0) return rc;
  if (field != 0 && field != 1) return fail(FDTD_ERR_ARG, "field must be 0 (E) or 1 (H)");
  if (d->x_wrap) return fail(FDTD_ERR_UNSUPPORTED, "periodic x boundary across slabs: drive the parts yourself");
  if ((rc = g_join.ensure()) != 0) return rc;
  return (d->dtype == FDTD_F32 ? sharded_halfstep<float, float>(d, h, field, q, slot, stream) : d->dtype == FDTD_F64 ? sharded_halfstep<double, double>(d, h, field, q, slot, stream) : sharded_halfstep<float, double>(d, h, field, q, slot, stream));
}

int fdtd_run_sharded(const fdtd_desc* d, fdtd_halo* h, int64_t q0, int64_t nsteps, int64_t slot0, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if ((rc = check_halo(d, h)) != 0) return rc;
  if (nsteps < 0) return fail(FDTD_ERR_ARG, "nsteps < 0");
  if (d->x_wrap) return fail(FDTD_ERR_UNSUPPORTED, "periodic x boundary across slabs: drive the parts yourself");
  if ((rc = g_join.ensure()) != 0) return rc;
  int64_t s = 0;
  if (nsteps >= 2 && fuse_eh_sharded(d, h)) {
    // pairs of temporally fused steps, A -> B -> A: the caller's buffers hold the result again.  Everything that is
    // still running on the side stream (the boundary plane of an earlier two-half-step step) comes first.
    if ((rc = stream_after((cudaStream_t)stream, (cudaStream_t)h->side_stream, g_join.to_main)) != 0) return rc;
    for (; s + 2 <= nsteps; s += 2) {
      for (int parity = 0; parity < 2; ++parity) {
        rc = d->dtype == FDTD_F32 ? fused_sharded_step<float>(d, h, parity, q0 + s + parity, slot0 + s + parity, stream)
                                  : fused_sharded_step<double>(d, h, parity, q0 + s + parity, slot0 + s + parity, stream);
        if (rc) return rc;
      }
    }
  }
  for (; s < nsteps; ++s) {
    for (int field = 0; field < 2; ++field) {
      rc = (d->dtype == FDTD_F32 ? sharded_halfstep<float, float>(d, h, field, q0 + s, slot0 + s, stream) : d->dtype == FDTD_F64 ? sharded_halfstep<double, double>(d, h, field, q0 + s, slot0 + s, stream) : sharded_halfstep<float, double>(d, h, field, q0 + s, slot0 + s, stream));
      if (rc) return rc;
    }
  }
  return FDTD_OK;
}

int fdtd_halo_refresh(const fdtd_desc* d, fdtd_halo* h, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if ((rc = check_halo(d, h)) != 0) return rc;
  for (int field = 0; field < 2; ++field) {
    const bool is_e = field == 0;
    if (is_e ? h->has_left : h->has_right) {
      rc = fdtd_halo_push(d, field, is_e ? h->left_ghost_y : h->right_ghost_y, is_e ? h->left_ghost_z : h->right_ghost_z,
                          stream);
      if (rc) return rc;
      rc = fdtd_halo_signal(is_e ? h->left_flag : h->right_flag, h->count[field] + 1, stream);
      if (rc) return rc;
    }
    h->count[field] += 1;
  }
  // E ghosts come from the right neighbour, H ghosts from the left one
  if (h->has_right) {
    rc = fdtd_halo_wait(h->flags, h->count[0], h->error, h->timeout_ns, stream);
    if (rc) return rc;
  }
  if (h->has_left) {
    rc = fdtd_halo_wait(h->flags + 1, h->count[1], h->error, h->timeout_ns, stream);
    if (rc) return rc;
  }
  return FDTD_OK;
}
#endif

#ifndef FDTD_EMU
// ---- CUDA-graph replay of step chunks (small, launch-bound grids) ---------------------------------
// One graph = FDTD_GRAPH_STEPS full steps with the fused kernels; the waveform index and ring slot of
// node s are s + dyn[], and dyn[] is set by a one-thread kernel before each replay.  Executable graphs
// are cached per descriptor content (every pointer and size they bake in), a few per host thread.
#define FDTD_GRAPH_STEPS 32
#define FDTD_GRAPH_CACHE 4     /* descriptors (grids) whose graphs are kept per host thread */
namespace {
struct GraphEntry {
  fdtd_desc desc;              // the exact descriptor the graph was captured from ...
  std::vector<fdtd_source> sources;       // ... and the source / detector tables it points to
  std::vector<fdtd_detector> detectors;
  cudaGraphExec_t exec = nullptr;
  uint64_t used = 0;           // LRU stamp
  bool same(const fdtd_desc* d) const {
    fdtd_desc a = desc, b = *d;
    a.sources = b.sources = nullptr;      // compared by content
    a.detectors = b.detectors = nullptr;
    if (memcmp(&a, &b, sizeof(fdtd_desc)) != 0) return false;
    if ((size_t)d->n_sources != sources.size() || (size_t)d->n_detectors != detectors.size()) return false;
    if (d->n_sources && memcmp(sources.data(), d->sources, sizeof(fdtd_source) * sources.size()) != 0) return false;
    return !d->n_detectors || memcmp(detectors.data(), d->detectors, sizeof(fdtd_detector) * detectors.size()) == 0;
  }
};
struct GraphCache {
  GraphEntry entry[FDTD_GRAPH_CACHE];
  uint64_t clock = 0;
  cudaStream_t capture_stream = nullptr;
};
thread_local GraphCache g_graph;

int graph_for(const fdtd_desc* d, cudaGraphExec_t* out) {
  // a graph bakes in every pointer, size and table of the descriptor: reuse it only for a byte-identical one
  GraphEntry* slot = &g_graph.entry[0];
  for (GraphEntry& e : g_graph.entry) {
    if (e.exec && e.same(d)) {
      e.used = ++g_graph.clock;
      *out = e.exec;
      return FDTD_OK;
    }
    if (e.used < slot->used) slot = &e;        // least recently used (or empty) entry
  }
  if (slot->exec) {
    cudaGraphExecDestroy(slot->exec);
    slot->exec = nullptr;
  }
  if (!g_graph.capture_stream &&
      cudaStreamCreateWithFlags(&g_graph.capture_stream, cudaStreamNonBlocking) != cudaSuccess)
    return fail(FDTD_ERR_CUDA, "cudaStreamCreate for graph capture failed");
  cudaStream_t cs = g_graph.capture_stream;
  if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    return fail(FDTD_ERR_CUDA, "cudaStreamBeginCapture failed");
  int rc = FDTD_OK;
  for (int64_t s = 0; s < FDTD_GRAPH_STEPS && rc == FDTD_OK; ++s) {
    rc = update_E_nocheck(d, 0, 0, cs, s);
    if (rc == FDTD_OK) rc = update_H_nocheck(d, 0, 0, cs, s);
  }
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(cs, &graph);
  if (rc != FDTD_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess || !graph) return fail(FDTD_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&slot->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    slot->exec = nullptr;
    return fail(FDTD_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  }
  memcpy(&slot->desc, d, sizeof(fdtd_desc));
  slot->sources.assign(d->sources, d->sources + d->n_sources);
  slot->detectors.assign(d->detectors, d->detectors + d->n_detectors);
  slot->used = ++g_graph.clock;
  *out = slot->exec;
  return FDTD_OK;
}
}  // namespace
#endif

int fdtd_fuse_eh_active(const fdtd_desc* d) {
  int rc = validate(d);
  if (rc) return rc;
  return fuse_eh_eligible(d) ? 1 : 0;
}

int fdtd_run(const fdtd_desc* d, int64_t q0, int64_t nsteps, int64_t slot0, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (nsteps < 0) return fail(FDTD_ERR_ARG, "nsteps < 0");
  if (d->Nx != d->Nx_global && nsteps > 0)
    return fail(FDTD_ERR_UNSUPPORTED, "fdtd_run on an x-sharded slab: drive the half-steps and the halo exchange per step");
  int64_t s = 0;
  {
    // pairs of temporally fused steps: A -> B -> A, so the caller's buffers hold the result again
    if (nsteps >= 2 && fuse_eh_eligible(d)) {
      for (; s + 2 <= nsteps; s += 2) {
        rc = d->dtype == FDTD_F32
                 ? fused_eh_step_unsharded<float>(d, d->E, d->E2, d->H, d->H2, q0 + s, slot0 + s, stream, 0)
                 : fused_eh_step_unsharded<double>(d, d->E, d->E2, d->H, d->H2, q0 + s, slot0 + s, stream, 0);
        if (rc) return rc;
        rc = d->dtype == FDTD_F32
                 ? fused_eh_step_unsharded<float>(d, d->E2, d->E, d->H2, d->H, q0 + s + 1, slot0 + s + 1, stream, 1)
                 : fused_eh_step_unsharded<double>(d, d->E2, d->E, d->H2, d->H, q0 + s + 1, slot0 + s + 1, stream, 1);
        if (rc) return rc;
      }
    }
  }
#ifndef FDTD_EMU
  if (s == 0 && d->use_graphs && d->dyn && post_is_fused(d) && nsteps >= FDTD_GRAPH_STEPS) {
    // the whole replay must stay inside every waveform table and detector ring
    int64_t chunks = nsteps / FDTD_GRAPH_STEPS;
    int64_t wave_base = 0;
    bool ok = true, have_src = false;
    for (int n = 0; n < d->n_sources; ++n) {
      const fdtd_source& S = d->sources[n];
      int64_t w = q0 - S.wave_q0;
      if (w < 0 || w + chunks * FDTD_GRAPH_STEPS > S.wave_len) ok = false;
      if (have_src && w != wave_base) ok = false;  // one base for all tables
      wave_base = w;
      have_src = true;
    }
    for (int n = 0; n < d->n_detectors; ++n)
      if (d->detectors[n].n > 0 && (slot0 < 0 || slot0 + chunks * FDTD_GRAPH_STEPS > d->detectors[n].capacity)) ok = false;
    if (ok) {
      cudaGraphExec_t exec = nullptr;
      rc = graph_for(d, &exec);
      if (rc) return rc;
      for (int64_t c = 0; c < chunks; ++c) {
        FDTD_LAUNCH((fdtd::set_dyn_kernel), dim3(1), dim3(1), stream, (i64*)d->dyn, (i64)(wave_base + s), (i64)(slot0 + s));
        rc = check_launch("set_dyn");
        if (rc) return rc;
        cudaError_t e = cudaGraphLaunch(exec, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(FDTD_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(e));
        g_launches.fetch_add(2 * FDTD_GRAPH_STEPS, std::memory_order_relaxed);
        s += FDTD_GRAPH_STEPS;
      }
    }
  }
#endif
  for (; s < nsteps; ++s) {
    rc = update_E_nocheck(d, q0 + s, slot0 + s, stream);
    if (rc) return rc;
    rc = update_H_nocheck(d, q0 + s, slot0 + s, stream);
    if (rc) return rc;
  }
  return FDTD_OK;
}

}  // extern "C"
