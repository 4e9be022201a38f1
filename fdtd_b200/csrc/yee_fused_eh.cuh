// yee_fused_eh.cuh -- single-pass E+H update of a whole homogeneous grid (temporal fusion).
//
// The two-half-step algorithm moves 18 words per cell and step (E half-step: read H, read E, write E; H half-step:
// read E, read H, write H).  Here a block marching along x updates E[i] and, one plane behind, H[i-1] in the same
// pass, so that every field word is read once and written once per STEP: 12 words.  What makes this legal:
//   * H[i-1]'s curl needs E_new of the y+1 / z+1 neighbours.  Inside the tile they come through shared memory;
//     at the tile edge a one-cell halo of E_new is recomputed by one extra row and one extra lane of threads;
//   * blocks are not synchronised with each other, so nothing may be updated in place: E and H are ping-pong
//     buffers (a fused step reads buffer A and writes buffer B; an even number of fused steps ends in the caller's
//     buffers, an odd remainder is run with the ordinary in-place kernels), and so is psi_E (halo cells recompute
//     E_new from the OLD psi_E while the owner writes the new one); psi_H is owner-only and stays in place;
//   * every CPML slab (registration order) and the masked one-sided differences on the six faces are handled inside
//     the kernel: one launch per step, no shell; E point sources are applied to every recomputed value.
// Arithmetic per cell is the same as in halfstep_kernel, operation by operation: results are bit-identical
// (tests/test_gpu_parity.py::test_temporally_fused_steps_equal_two_half_steps).
//
// History (profiles/README.md): round 1 had three variants -- (1) shared-memory exchange with direct
// global loads on the CPML-free interior plus twelve shell launches of the ordinary kernel, (2) a register-tiled kernel
// without any inter-thread communication, (3) this one with per-thread cp.async staging.  On the B200 at 1024^3 f32
// they ran 12.74 / 15.0 / 11.69 ms per step against 12.62 for the two half-steps; (1) and (2) were deleted.  Since then:
// TMA staging (10.77), psi of the block's z slab staged by bulk copies + z+1 neighbour by shuffle (10.20), CPML tables in
// shared memory (9.98); the cp.async staging, a split mbarrier barrier and several prefetch schemes were measured and
// removed.  The kernel is DRAM-bound at 5.7 TB/s (profiles/r2_s15/).
#pragma once

namespace fdtd {

#ifndef FDTD_FUSED_MAX_CHUNKS
#define FDTD_FUSED_MAX_CHUNKS 64     // x-chunks per launch (kernel parameter space)
#endif

template <typename T>
struct FusedParams {
  int Ny, Nz;
  i64 plane;
  int x0, x1, y0, y1, z0, z1;  // interior box (cells); z0, z1 multiples of the vector width
  // the march along x is cut into chunks [xstart[k], xstart[k+1]); block z handles chunk chunk0 + z * chunk_step (a launch
  // may cover a subset of the chunks)
  int chunk0, chunk_step;
  int xstart[FDTD_FUSED_MAX_CHUNKS + 1];
  const T* Ein[3];
  T* Eout[3];
  const T* Hin[3];
  T* Hout[3];
  T ce[3], ch[3];  // sc * background eps^-1 / mu^-1, rounded as the reference rounds them
  int n_src;
  SrcK<T> src[FDTD_FUSED_MAX];  // soft point-list sources on E (ascending idx)
  // the box is the WHOLE grid: every CPML slab and the masked one-sided differences at the six faces are handled
  // inside the kernel, in registration order.  psi_E is ping-pong like the fields, psi_H in place.
  int Nx;          // first x index without a plane of its own in Eout: the local extent, + 1 when a right neighbour's
                   // first plane lies in the ghost plane behind it
  // x-sharded slabs: global index of local plane 0 and the global extent (the face masks are global), the peer's ghost
  // planes the y / z components of E_new[plane 0] are stored into as well (or null), and whether the H update of the
  // LAST local plane is left to a separate launch (it needs the right neighbour's E_new, which arrives meanwhile)
  int x_offset, Nx_global;
  T* push_y;
  T* push_z;
  int skip_last_h;
  int psi_stage;   // the psi arrays of every z slab are 16-byte aligned: a block may stage them with bulk copies
  int n_sl;
  struct Slab {
    int axis, lo, t, lo_al, tp;   // lo: LOCAL coordinate of the slab's first cell along its axis (x slabs: may be
                                  // negative on a slab that starts inside it); lo_al / tp: padded psi rows of z slabs
    int xs, xe;                   // x slabs: local planes [xs, xe) have psi storage
    i64 count;
    const T* psiE_in;
    T* psiE_out;
    T* psiH;
    const T* bE;
    const T* cE;
    const T* bH;
    const T* cH;
  } sl[6];
};

// soft point sources of the box on the VEC recomputed E values of a thread, registration order
// (fdtd/sources.py:93-109, 278-297); inlined: an out-of-line call would force the field vectors into local memory
template <typename T, int VEC>
FDTD_DEV void fused_sources_vec(const FusedParams<T>& P, int i, int j, int k0, i64 off, Pack<T, VEC>& e0,
                                    Pack<T, VEC>& e1, Pack<T, VEC>& e2) {
  for (int s = 0; s < P.n_src; ++s) {
    const SrcK<T>& S = P.src[s];
    if (i < S.bb[0] || i >= S.bb[1] || j < S.bb[2] || j >= S.bb[3] || k0 + VEC <= S.bb[4] || k0 >= S.bb[5]) continue;
    const T wv = S.wave[S.w];
    for (int n = lower_bound_i64(S.idx, S.n, off); n < S.n && S.idx[n] < off + VEC; ++n) {
      const int de = (int)(S.idx[n] - off);
      const T v = S.profile[n] * wv;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (e == de) {
          if (S.comp == 0) e0.v[e] = e0.v[e] + v;
          else if (S.comp == 1) e1.v[e] = e1.v[e] + v;
          else e2.v[e] = e2.v[e] + v;
        }
      }
    }
  }
}

// core tile of a block: R rows x L vector lanes, plus one halo row and one halo lane of threads.  7 x 31 + halo =
// 8 full warps; two blocks per SM.  Measured at 1024^3 f32 (ms per step): 4x32 (165 threads, 3 blocks) 14.3,
// 4x31 (3 blocks) 12.8, 8x16 (2) 13.1, 15x31 (1) 12.1, 3x31 (4) 11.9, 7x31 (2) 11.7   profiles/r2_fused_variants.txt
#ifndef FDTD_FUSED_ROWS
#define FDTD_FUSED_ROWS 7
#endif
#ifndef FDTD_FUSED_LANES
#define FDTD_FUSED_LANES 31
#endif
constexpr int FUSED_R = FDTD_FUSED_ROWS;   // core rows per block
constexpr int FUSED_L = FDTD_FUSED_LANES;  // core vector lanes per block

// ---- the inputs of each plane are staged in shared memory two planes ahead -----------------------------------------
// No thread ever waits for a global load it has just issued: every plane's H_old tile (own cells + the y-1 row and
// the z-1 vector) and E_old tile arrive in one of three shared-memory stages through bulk copies issued two
// iterations before they are consumed.  Each global word is requested once per block (the y-1 / z-1 neighbours come
// out of the staged tile).
// Iteration i:  wait for the stage of plane i -> barrier (everyone's E_new[i-1] is published, stage (i-1)%3 is free)
// -> issue the copies of plane i+2 into stage (i-1)%3 -> E_new[i] from stage i%3 -> H_new[i-1] from the published
// E_new[i-1] -> publish E_new[i].
#ifdef FDTD_EMU
#define FDTD_FFS(x) __builtin_ffs((int)(x))
#define FDTD_DYN_SMEM(name) alignas(128) static unsigned char name[256 << 10]
#else
#define FDTD_FFS(x) __ffs((int)(x))
#define FDTD_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif
#ifndef FDTD_FUSED_PIPE_MIN_BLOCKS
#define FDTD_FUSED_PIPE_MIN_BLOCKS 2
#endif

// ---- staging by TMA: one elected thread issues six bulk tensor copies per plane (the H_old and E_old tiles of the
// three components, 3-D boxes with hardware zero fill outside the grid), completion is counted on one mbarrier per
// stage that every thread of the block waits on.  No per-thread copy instructions, address arithmetic or predicates.
#ifdef FDTD_EMU
template <typename T>
struct TmaMap {           // (CPU interpreter) the tensor a map describes: [n2][n1][n0] elements at base
  const T* base;
  int n0, n1, n2;
};
#define FDTD_MBAR_INIT(bar) emu::mbar_init(bar)
#define FDTD_MBAR_EXPECT(bar, bytes) ((void)(bytes))
#define FDTD_MBAR_WAIT(bar, parity) emu::mbar_wait(bar)
#define FDTD_TMA_LOAD_3D(dst, map, c0, c1, c2, b0, b1, bar) \
  emu::tma_load_3d((dst), (map)->base, sizeof(*(map)->base), (map)->n0, (map)->n1, (map)->n2, (c0), (c1), (c2), (b0), (b1), (bar))
#define FDTD_BULK_LOAD(dst, src, bytes, bar) emu::bulk_load((dst), (src), (bytes), (bar))
#define FDTD_SHFL_DOWN1(v) emu::shfl_down1(v)
#else
template <typename T>
struct TmaMap {
  CUtensorMap m;
};
FDTD_DEV unsigned fused_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#define FDTD_MBAR_INIT(bar) \
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fdtd::fused_smem_addr(bar)) : "memory")
#define FDTD_MBAR_EXPECT(bar, bytes)                                                                      \
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fdtd::fused_smem_addr(bar)), \
               "r"((unsigned)(bytes))                                                                      \
               : "memory")
#define FDTD_MBAR_WAIT(bar, parity)                                                                        \
  asm volatile(                                                                                            \
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\n" \
      "bra LAB_WAIT;\nDONE:\n}" ::"r"(fdtd::fused_smem_addr(bar)),                                          \
      "r"((unsigned)(parity))                                                                              \
      : "memory")
#define FDTD_TMA_LOAD_3D(dst, map, c0, c1, c2, b0, b1, bar)                                                       \
  asm volatile(                                                                                                   \
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"( \
          fdtd::fused_smem_addr(dst)),                                                                            \
      "l"((unsigned long long)(map)), "r"((int)(c0)), "r"((int)(c1)), "r"((int)(c2)), "r"(fdtd::fused_smem_addr(bar)) \
      : "memory")
// a contiguous run of global memory (16-byte aligned, a multiple of 16 bytes) into shared memory, counted on `bar`
#define FDTD_BULK_LOAD(dst, src, bytes, bar)                                                                     \
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(  \
                   fdtd::fused_smem_addr(dst)),                                                                  \
               "l"(src), "r"((unsigned)(bytes)), "r"(fdtd::fused_smem_addr(bar))                                 \
               : "memory")
#define FDTD_SHFL_DOWN1(v) __shfl_down_sync(0xffffffffu, (v), 1)
#endif

template <typename T>
struct FusedTmaMaps {
  TmaMap<T> h[3], e[3];   // the three components of the input H and E (ghost x-planes included: plane i is x = i + 1)
};
#ifndef FDTD_FUSED_EARLY_XPSI
#define FDTD_FUSED_EARLY_XPSI 1          // psi of x-slab planes is loaded at the top of the iteration (see the kernel)
#endif

#ifndef FDTD_FUSED_SHFL
#define FDTD_FUSED_SHFL 1        // a tile row is one warp (31 core lanes + the halo lane): the z+1 neighbour's E_new comes from
                                 // the next lane by shuffle, only the two components differenced along y (Ex, Ez) are published
                                 // through shared memory
#endif
#ifndef FDTD_FUSED_PSI_STAGE
#define FDTD_FUSED_PSI_STAGE 1   // psi of the z slab a block touches (two of every nine blocks at 1024^3) is staged with the
                                 // plane's field tiles: the rows of a tile are ONE contiguous run of a z slab's psi array, so
                                 // four 1-D bulk copies per plane take the only global loads off the per-plane critical path
#endif

#ifndef FDTD_FUSED_TAB_T
#define FDTD_FUSED_TAB_T 32      // thickest CPML slab the fused kernel takes (cells); thicker ones: two half-steps
#endif

template <typename T, int VEC>
struct FusedPipeLayout {
  static constexpr int R = FUSED_R, L = FUSED_L;
  static constexpr bool SHFL = (FDTD_FUSED_SHFL != 0) && (L + 1 == 32);
  static constexpr int HV = L + 2, EV = L + 1;                 // vectors per staged row
  static constexpr int PADW = 128 / (int)sizeof(T);            // a TMA destination is 128-byte aligned
  static constexpr int HC = ((R + 2) * HV * VEC + PADW - 1) / PADW * PADW;   // words per staged H component
  static constexpr int EC = ((R + 1) * EV * VEC + PADW - 1) / PADW * PADW;   // words per staged E component
  static constexpr int H_WORDS = 3 * HC;                       // H_old tile: rows j0-1 .. j0+R, vectors -1 .. L
  static constexpr int E_WORDS = 3 * EC;                       // E_old tile: rows j0 .. j0+R, vectors 0 .. L
  static constexpr int FIELD_WORDS = H_WORDS + E_WORDS;
  // staged psi of one z slab: psi_E[0], psi_E[1] (rows j0 .. j0+R of plane i), psi_H[0], psi_H[1] (rows j0 .. j0+R-1 of
  // plane i-1), each (rows x tp) words as they lie in memory, tp <= PSI_TP
  static constexpr int PSI_TP = 16;
  static constexpr int PSI_ARR = (R + 1) * PSI_TP;
  static constexpr int PSI_WORDS = FDTD_FUSED_PSI_STAGE ? 4 * PSI_ARR : 0;
  static constexpr int STAGE_WORDS = FIELD_WORDS + PSI_WORDS;
  static constexpr int XC = (R + 1) * EV * VEC;                // words per published component
  static constexpr int X_COMPS = SHFL ? 2 : 3;                 // published components of E_new (SHFL: Ex, Ez)
  static constexpr int X_WORDS = X_COMPS * XC;                 // one published E_new tile
  static constexpr unsigned TMA_BYTES = 3u * ((R + 2) * HV + (R + 1) * EV) * VEC * (unsigned)sizeof(T);  // per plane
  static constexpr int STAGES = 3;
  // the CPML coefficient tables (b_E, c_E, b_H, c_H of every slab, TAB_T entries each) live in shared memory: their
  // loads sat on the long scoreboard in every slab cell (profiles/r2_s13/stalls_fused_v2.txt)
  static constexpr int TAB_T = FDTD_FUSED_TAB_T;
  static constexpr int TAB_WORDS = 6 * 4 * TAB_T;
  static constexpr size_t TAB_OFFSET = sizeof(T) * (size_t)(STAGES * STAGE_WORDS + 2 * X_WORDS);
  static constexpr size_t BAR_OFFSET = TAB_OFFSET + sizeof(T) * (size_t)TAB_WORDS;   // the mbarriers
  static constexpr size_t BYTES = BAR_OFFSET + 8 * STAGES;
};

// CPML update of one slab (axis AX) for the VEC cells of a thread, psi read from `psi_in` and (if `store`) written
// to `psi_out` -- the same operations as slab_cells (yee_kernels.cuh).  With (AX, U, W) cyclic: psi[0] is driven by
// the one-sided difference of G_W along AX and corrects F_U with sign -, psi[1] by that of G_U and corrects F_W with
// sign +.  Everything is selected at compile time, so the field vectors stay in registers.
template <typename T, int VEC>
struct FusedDiffs {
  T zy[VEC], yz[VEC], xz[VEC], zx[VEC], yx[VEC], xy[VEC];   // d G_c / d a, named "ca"
};

template <typename T, int VEC, bool IS_E, int AX>
FDTD_DEV void fused_slab_cells(const typename FusedParams<T>::Slab& S, Pack<T, VEC>& a, Pack<T, VEC>& b, int l0,
                               const FusedDiffs<T, VEC>& D, Pack<T, VEC>& f0, Pack<T, VEC>& f1, Pack<T, VEC>& f2,
                               const T (&coef)[3], const T* bt, const T* ct) {
  constexpr int U = (AX + 1) % 3, W = (AX + 2) % 3;
  Pack<T, VEC>& fu = U == 0 ? f0 : (U == 1 ? f1 : f2);
  Pack<T, VEC>& fw = W == 0 ? f0 : (W == 1 ? f1 : f2);
  const T (&d0)[VEC] = AX == 0 ? D.zx : (AX == 1 ? D.xy : D.yz);
  const T (&d1)[VEC] = AX == 0 ? D.yx : (AX == 1 ? D.zy : D.xz);
  const T cu = coef[U], cw = coef[W];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const int ll = AX == 2 ? l0 + e : l0;
    if (ll >= 0 && ll < S.t) {
      const T bb = bt[ll];
      const T cc = ct[ll];
      T p0 = a.v[e] * bb;
      T p1 = b.v[e] * bb;
      if (IS_E ? (ll >= 1) : (ll < S.t - 1)) {   // fdtd/boundaries.py:439-454, 467-482
        p0 = p0 + d0[e] * cc;
        p1 = p1 + d1[e] * cc;
      }
      a.v[e] = p0;
      b.v[e] = p1;
      const T phi_u = T(0) - p0;
      const T phi_w = p1 - T(0);
      if (IS_E) {
        fu.v[e] = fu.v[e] + cu * phi_u;
        fw.v[e] = fw.v[e] + cw * phi_w;
      } else {
        fu.v[e] = fu.v[e] - cu * phi_u;
        fw.v[e] = fw.v[e] - cw * phi_w;
      }
    }
  }
}

// One slab's CPML update of a thread's VEC cells: psi comes from global memory at `idx`, or -- the z slab a block
// stages -- from the shared-memory copy at `sp`; the new psi goes to global memory (`store`: owner threads only).
template <typename T, int VEC, bool IS_E>
FDTD_DEV void fused_slab_update(const typename FusedParams<T>::Slab& S, const T* psi_in, T* psi_out, bool store,
                                const T* sp, int sp_pitch, i64 idx, int l0, const FusedDiffs<T, VEC>& D,
                                Pack<T, VEC>& f0, Pack<T, VEC>& f1, Pack<T, VEC>& f2, const T (&coef)[3],
                                const T* tab, bool early, const Pack<T, VEC>& ea, const Pack<T, VEC>& eb) {
  const T* bt = tab + (IS_E ? 0 : 2) * FDTD_FUSED_TAB_T;   // (shared memory: b_E, c_E, b_H, c_H of this slab)
  const T* ct = bt + FDTD_FUSED_TAB_T;
  Pack<T, VEC> a, b;
  if (early) {             // loaded at the top of the iteration (x slabs)
    a = ea;
    b = eb;
  } else if (sp != nullptr) {
    a = ldv<T, VEC>(sp);
    b = ldv<T, VEC>(sp + sp_pitch);
  } else {
    a = ldv<T, VEC>(psi_in + idx);
    b = ldv<T, VEC>(psi_in + S.count + idx);
  }
  if (S.axis == 0) fused_slab_cells<T, VEC, IS_E, 0>(S, a, b, l0, D, f0, f1, f2, coef, bt, ct);
  else if (S.axis == 1) fused_slab_cells<T, VEC, IS_E, 1>(S, a, b, l0, D, f0, f1, f2, coef, bt, ct);
  else fused_slab_cells<T, VEC, IS_E, 2>(S, a, b, l0, D, f0, f1, f2, coef, bt, ct);
  if (store) {
    stv<T, VEC>(psi_out + idx, a);
    stv<T, VEC>(psi_out + S.count + idx, b);
  }
}

// one plane of the block's tile into a stage, by TMA (thread 0 only); zs >= 0: also the psi rows of z slab zs --
// psi_E of plane ip (rows j0 .. j0+R) and psi_H of plane ip-1 (rows j0 .. j0+R-1; with_h), contiguous in memory
template <typename T, int VEC>
FDTD_DEV void fused_stage_tma(const FusedParams<T>& P, const FusedTmaMaps<T>& M, T* stage, unsigned long long* bar,
                              int kz0, int j0, int ip, int zs, bool with_h) {
  using Lay = FusedPipeLayout<T, VEC>;
  unsigned bytes = Lay::TMA_BYTES, be = 0, bh = 0;
  if (FDTD_FUSED_PSI_STAGE && zs >= 0) {
    const int rows_e = P.Ny - j0 < Lay::R + 1 ? P.Ny - j0 : Lay::R + 1;
    const int rows_h = P.Ny - j0 < Lay::R ? P.Ny - j0 : Lay::R;
    be = (unsigned)(rows_e * P.sl[zs].tp) * (unsigned)sizeof(T);
    bh = with_h ? (unsigned)(rows_h * P.sl[zs].tp) * (unsigned)sizeof(T) : 0u;
    bytes += 2 * (be + bh);
  }
  FDTD_MBAR_EXPECT(bar, bytes);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    FDTD_TMA_LOAD_3D(stage + c * Lay::HC, &M.h[c], kz0 - VEC, j0 - 1, ip + 1, Lay::HV * VEC, Lay::R + 2, bar);
    FDTD_TMA_LOAD_3D(stage + Lay::H_WORDS + c * Lay::EC, &M.e[c], kz0, j0, ip + 1, Lay::EV * VEC, Lay::R + 1, bar);
  }
  if (FDTD_FUSED_PSI_STAGE && zs >= 0) {
    const typename FusedParams<T>::Slab& S = P.sl[zs];
    T* pt = stage + Lay::FIELD_WORDS;
    const T* se = S.psiE_in + ((i64)ip * P.Ny + j0) * S.tp;
    FDTD_BULK_LOAD(pt, se, be, bar);
    FDTD_BULK_LOAD(pt + Lay::PSI_ARR, se + S.count, be, bar);
    if (with_h) {
      const T* sh = S.psiH + ((i64)(ip - 1) * P.Ny + j0) * S.tp;
      FDTD_BULK_LOAD(pt + 2 * Lay::PSI_ARR, sh, bh, bar);
      FDTD_BULK_LOAD(pt + 3 * Lay::PSI_ARR, sh + S.count, bh, bar);
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__((FUSED_R + 1) * (FUSED_L + 1), FDTD_FUSED_PIPE_MIN_BLOCKS)
    fused_eh_pipe_kernel(const __grid_constant__ FusedParams<T> P, const __grid_constant__ FusedTmaMaps<T> M) {
  using Lay = FusedPipeLayout<T, VEC>;
  constexpr int R = Lay::R, L = Lay::L, HV = Lay::HV, EV = Lay::EV;
  static_assert(sizeof(T) * VEC == 16, "the staged copies are 16 bytes wide");
  FDTD_DYN_SMEM(smem_raw);
  T* const stages = reinterpret_cast<T*>(smem_raw);
  T* const xch = stages + Lay::STAGES * Lay::STAGE_WORDS;
  T* const tabs = reinterpret_cast<T*>(smem_raw + Lay::TAB_OFFSET);
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem_raw + Lay::BAR_OFFSET);

  const int tid = threadIdx.x;
  const int r = tid / (L + 1), l = tid % (L + 1);
  const int j0 = P.y0 + blockIdx.y * R, kz0 = P.z0 + blockIdx.x * L * VEC;
  const int j = j0 + r, k0 = kz0 + l * VEC;
  const int Nz = P.Nz;
  const bool core = (r < R) && (l < L) && (j < P.y1) && (k0 < P.z1);
  // (y1 == Ny / z1 == Nz: there is no y+1 neighbour row / z+1 neighbour column)
  const bool active = (j <= P.y1) && (k0 <= P.z1) && (j < P.Ny) && (k0 < Nz);
  const bool inside = (j < P.y1) && (k0 < P.z1);
  const i64 plane = P.plane;
  const i64 p = (i64)j * Nz + k0;
  const int xa = P.xstart[P.chunk0 + (int)blockIdx.z * P.chunk_step];
  const int xb = P.xstart[P.chunk0 + (int)blockIdx.z * P.chunk_step + 1];

  bool src_yz = false;
  for (int s = 0; s < P.n_src; ++s)
    src_yz |= (j >= P.src[s].bb[2]) && (j < P.src[s].bb[3]) && (k0 + VEC > P.src[s].bb[4]) && (k0 < P.src[s].bb[5]);
  // CPML slabs this thread's cells lie in (loop-invariant)
  unsigned sl_hit = 0;   // bit s: this thread's cells lie in y / z slab s (x slabs are decided per plane)
  unsigned xs_bits = 0;  // bit s: slab s is an x slab
  for (int s = 0; s < P.n_sl; ++s) {
    const bool hit = P.sl[s].axis == 1 ? (j >= P.sl[s].lo && j < P.sl[s].lo + P.sl[s].t)
                                       : (k0 - P.sl[s].lo + VEC > 0) && (k0 - P.sl[s].lo < P.sl[s].t);
    sl_hit |= (hit && P.sl[s].axis != 0) ? (1u << s) : 0u;
    // (x slabs that do not reach into this block's planes [xa, xb] are never looked at again)
    xs_bits |= (P.sl[s].axis == 0 && P.sl[s].xe > xa && P.sl[s].xs <= xb) ? (1u << s) : 0u;
  }
  unsigned hit_prev = 0;   // slabs (x slabs included) the cells of plane i-1 lie in: what the H update of i-1 needs
  // the z slab whose psi this block stages in shared memory (block-uniform): the first one a lane of the tile lies in
  int zs = -1;
  if (FDTD_FUSED_PSI_STAGE && P.psi_stage) {
    for (int s = P.n_sl - 1; s >= 0; --s)
      if (P.sl[s].axis == 2 && P.sl[s].tp <= Lay::PSI_TP && kz0 + (L + 1) * VEC > P.sl[s].lo &&
          kz0 < P.sl[s].lo + P.sl[s].t)
        zs = s;
  }
  const int zs_off = zs >= 0 ? r * P.sl[zs].tp + (k0 - P.sl[zs].lo_al) : 0;   // this thread's cells in a staged psi array

  if (tid == 0) {
    for (int s = 0; s < Lay::STAGES; ++s) FDTD_MBAR_INIT(bars + s);
  }
#ifndef FDTD_EMU
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  // the coefficient tables of every slab (thickness <= TAB_T: the host's eligibility test)
  for (int n = tid; n < P.n_sl * 4 * Lay::TAB_T; n += (R + 1) * (L + 1)) {
    const int s = n / (4 * Lay::TAB_T), w = (n / Lay::TAB_T) & 3, ll = n % Lay::TAB_T;
    const T* src = w == 0 ? P.sl[s].bE : (w == 1 ? P.sl[s].cE : (w == 2 ? P.sl[s].bH : P.sl[s].cH));
    tabs[n] = ll < P.sl[s].t ? src[ll] : T(0);
  }
  __syncthreads();
  // two planes in flight before the first one is consumed
  for (int s = 0; s < 2; ++s) {
    const int ip = xa + s;
    if (tid == 0 && ip <= xb && ip < P.x1)
      fused_stage_tma<T, VEC>(P, M, stages + (ip % 3) * Lay::STAGE_WORDS, bars + (ip % 3), kz0, j0, ip, zs, ip > xa);
  }

  Pack<T, VEC> hp0, hp1, hp2, ep0 = {}, ep1 = {}, ep2;
  if (active && inside && xa + P.x_offset > 0) {
    const i64 o = (i64)(xa - 1) * plane + p;
    hp0 = ldv<T, VEC>(P.Hin[0] + o);
    hp1 = ldv<T, VEC>(P.Hin[1] + o);
    hp2 = ldv<T, VEC>(P.Hin[2] + o);
  }

  const int xl = (P.skip_last_h && xb == P.x1) ? xb - 1 : xb;   // (see FusedParams::skip_last_h)
  int sx_prev = -1;        // the x slab whose psi_H of plane i-1 is loaded early (see below)
  for (int i = xa; i <= xl; ++i) {
    unsigned hit_now = sl_hit;   // + the x slabs plane i lies in (the same for the whole block)
    int sx = -1;                 // the first of them
    for (unsigned m = xs_bits; m != 0; m &= m - 1) {
      const int s = FDTD_FFS(m) - 1;
      const bool in = i >= P.sl[s].xs && i < P.sl[s].xe;
      hit_now |= in ? (1u << s) : 0u;
      if (in && sx < 0) sx = s;
    }
    // Inside an x slab EVERY thread needs psi, and its loads were the only global loads on the per-plane critical path
    // (an x-PML plane cost 2.3 ordinary planes, profiles/r2_s16/): they are issued here, ahead of the wait for the
    // staged plane, the barrier and the curls -- psi_E of plane i and psi_H of plane i-1.
    Pack<T, VEC> xea = {}, xeb = {}, xha = {}, xhb = {};
    const bool early_e = FDTD_FUSED_EARLY_XPSI && sx >= 0 && active && inside && i < P.x1;
    const bool early_h = FDTD_FUSED_EARLY_XPSI && sx_prev >= 0 && core && i > xa;
    if (early_e) {
      const i64 idx = (i64)(i - P.sl[sx].xs) * plane + p;
      xea = ldv<T, VEC>(P.sl[sx].psiE_in + idx);
      xeb = ldv<T, VEC>(P.sl[sx].psiE_in + P.sl[sx].count + idx);
    }
    if (early_h) {
      const i64 idx = (i64)(i - 1 - P.sl[sx_prev].xs) * plane + p;
      xha = ldv<T, VEC>(P.sl[sx_prev].psiH + idx);
      xhb = ldv<T, VEC>(P.sl[sx_prev].psiH + P.sl[sx_prev].count + idx);
    }
    // the k-th use of a stage's barrier completes phase k: plane i is use (i - xa) / 3 of stage i % 3
    if (i < P.x1) FDTD_MBAR_WAIT(bars + (i % 3), ((i - xa) / 3) & 1);
    // (tried: the barrier split into an mbarrier arrival after the publish below and a wait between the next E and H
    // updates, so that warps may drift by one E update: 10.15 instead of 10.00 ms per step, profiles/r2_s14/; an L2
    // prefetch of the tiles 3 / 4 / 6 planes ahead by cp.async.bulk.prefetch.tensor: 10.80 / 11.19 / 12.27 ms against
    // 9.98, profiles/r2_s15/ -- the kernel waits for DRAM at 5.7 TB/s, more requests in flight only queue longer;
    // evict-first stores (st.global.cs) of E_new / H_new / psi: 10.09 against 10.00, DRAM reads 29.77 instead of 29.97 GB,
    // profiles/r2_s18/)
    __syncthreads();
    {
      const int ip = i + 2;
      if (tid == 0 && ip <= xb && ip < P.x1)
        fused_stage_tma<T, VEC>(P, M, stages + (ip % 3) * Lay::STAGE_WORDS, bars + (ip % 3), kz0, j0, ip, zs, ip > xa);
    }
    const T* sH = stages + (i % 3) * Lay::STAGE_WORDS;
    const T* sE = sH + Lay::H_WORDS;
    const i64 off = (i64)i * plane + p;
    Pack<T, VEC> e0, e1, e2, h0, h1, h2;
    if (active) {
      if (inside && i < P.x1) {
        // ---- E_new[i] = E_old + (sc eps^-1) * curl_H(H_old)      (fdtd/grid.py:54-76, 283)
        const T* hrow = sH + ((r + 1) * HV + (l + 1)) * VEC;            // own vector of component 0
        constexpr int HC = Lay::HC;                                     // words per staged H component
        h0 = ldv<T, VEC>(hrow);
        h1 = ldv<T, VEC>(hrow + HC);
        h2 = ldv<T, VEC>(hrow + 2 * HC);
        const Pack<T, VEC> y0v = ldv<T, VEC>(hrow - HV * VEC);
        const Pack<T, VEC> y2v = ldv<T, VEC>(hrow + 2 * HC - HV * VEC);
        const T zs0 = hrow[-1];
        const T zs1 = hrow[HC - 1];
        const T* erow = sE + (r * EV + l) * VEC;
        constexpr int EC = Lay::EC;
        e0 = ldv<T, VEC>(erow);
        e1 = ldv<T, VEC>(erow + EC);
        e2 = ldv<T, VEC>(erow + 2 * EC);
        FusedDiffs<T, VEC> D;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const T zn0 = e == 0 ? zs0 : h0.v[e > 0 ? e - 1 : 0];
          const T zn1 = e == 0 ? zs1 : h1.v[e > 0 ? e - 1 : 0];
          // backward differences are masked on the faces i = 0, j = 0 and k = 0 (fdtd/grid.py:66-74)
          const bool face_x = i + P.x_offset == 0;
          const bool face_y = j == 0;
          const bool face_z = (e == 0) && (k0 == 0);
          const T d_zy = face_y ? T(0) : h2.v[e] - y2v.v[e];
          const T d_xy = face_y ? T(0) : h0.v[e] - y0v.v[e];
          const T d_yz = face_z ? T(0) : h1.v[e] - zn1;
          const T d_xz = face_z ? T(0) : h0.v[e] - zn0;
          const T d_zx = face_x ? T(0) : h2.v[e] - hp2.v[e];
          const T d_yx = face_x ? T(0) : h1.v[e] - hp1.v[e];
          D.yz[e] = d_yz;
          D.xz[e] = d_xz;
          D.zy[e] = d_zy;
          D.xy[e] = d_xy;
          D.zx[e] = d_zx;
          D.yx[e] = d_yx;
          e0.v[e] = e0.v[e] + P.ce[0] * (d_zy - d_yz);
          e1.v[e] = e1.v[e] + P.ce[1] * (d_xz - d_zx);
          e2.v[e] = e2.v[e] + P.ce[2] * (d_yx - d_xy);
        }
        // CPML slabs, registration order: psi_E from the OLD buffer; an x slab corrects Ey (-) and Ez (+), a y slab
        // Ez (-) and Ex (+), a z slab Ex (-) and Ey (+)      (fdtd/boundaries.py:433-459, 409-419)
        // (the interior -- no slab at all -- skips the loop with one test)
        for (int s = 0; hit_now != 0 && s < P.n_sl; ++s) {
          if (!((hit_now >> s) & 1u)) continue;
          const typename FusedParams<T>::Slab& S = P.sl[s];
          const bool store = core && i < xb;
          const T* sp = (FDTD_FUSED_PSI_STAGE && s == zs) ? sH + Lay::FIELD_WORDS + zs_off : nullptr;
          const i64 idx = S.axis == 0 ? (i64)(i - S.xs) * plane + p
                                      : (S.axis == 1 ? ((i64)i * S.t + (j - S.lo)) * Nz + k0
                                                     : ((i64)i * P.Ny + j) * S.tp + (k0 - S.lo_al));
          const int l0 = S.axis == 0 ? i - S.lo : (S.axis == 1 ? j - S.lo : k0 - S.lo);
          fused_slab_update<T, VEC, true>(S, S.psiE_in, S.psiE_out, store, sp, Lay::PSI_ARR, idx, l0, D, e0, e1, e2, P.ce,
                                          tabs + s * 4 * Lay::TAB_T, early_e && s == sx, xea, xeb);
        }
        if (src_yz) fused_sources_vec<T, VEC>(P, i, j, k0, off, e0, e1, e2);
        if (core && i < xb) {
          stv<T, VEC>(P.Eout[0] + off, e0);
          stv<T, VEC>(P.Eout[1] + off, e1);
          stv<T, VEC>(P.Eout[2] + off, e2);
          if (i == 0 && P.push_y != nullptr) {
            // the slab's boundary plane also goes into the left neighbour's ghost plane, over NVLink
            stv<T, VEC>(P.push_y + p, e1);
            stv<T, VEC>(P.push_z + p, e2);
          }
        }
      } else if (i < P.Nx) {
        // a shell cell (outside the box in y / z, or the plane x1): its E_new is already in memory
        e0 = ldv<T, VEC>(P.Eout[0] + off);
        e1 = ldv<T, VEC>(P.Eout[1] + off);
        e2 = ldv<T, VEC>(P.Eout[2] + off);
      }
    }

    // ---- H_new[i-1] = H_old - (sc mu^-1) * curl_E(E_new)      (fdtd/grid.py:29-51, 309)
    constexpr int XC = Lay::XC;                                         // words per published component
    constexpr int XZ = (Lay::X_COMPS - 1) * XC;                         // where the published Ez starts
    T nx0 = T(0), ny0 = T(0);   // Ex, Ey of the z+1 neighbour cell of this thread's last cell (E_new[i-1])
    if (Lay::SHFL) {
      nx0 = FDTD_SHFL_DOWN1(ep0.v[0]);
      ny0 = FDTD_SHFL_DOWN1(ep1.v[0]);
    }
    if (core && i > xa) {
      const T* x = xch + ((i - 1) & 1) * Lay::X_WORDS + (r * EV + l) * VEC;
      if (!Lay::SHFL) {
        nx0 = x[VEC];
        ny0 = x[XC + VEC];
      }
      Pack<T, VEC> hx = hp0, hy = hp1, hz = hp2;
      FusedDiffs<T, VEC> D;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const T ex_y = x[EV * VEC + e];
        const T ez_y = x[XZ + EV * VEC + e];
        const T ex_z = e == VEC - 1 ? nx0 : ep0.v[e < VEC - 1 ? e + 1 : 0];
        const T ey_z = e == VEC - 1 ? ny0 : ep1.v[e < VEC - 1 ? e + 1 : 0];
        // forward differences are masked on the faces i = Nx-1, j = Ny-1 and k = Nz-1 (fdtd/grid.py:41-49)
        const bool face_x = i + P.x_offset == P.Nx_global;
        const bool face_y = j == P.Ny - 1;
        const bool face_z = (e == VEC - 1) && (k0 + VEC == Nz);
        const T d_zy = face_y ? T(0) : ez_y - ep2.v[e];
        const T d_xy = face_y ? T(0) : ex_y - ep0.v[e];
        const T d_yz = face_z ? T(0) : ey_z - ep1.v[e];
        const T d_xz = face_z ? T(0) : ex_z - ep0.v[e];
        const T d_zx = face_x ? T(0) : e2.v[e] - ep2.v[e];
        const T d_yx = face_x ? T(0) : e1.v[e] - ep1.v[e];
        D.yz[e] = d_yz;
        D.xz[e] = d_xz;
        D.zy[e] = d_zy;
        D.xy[e] = d_xy;
        D.zx[e] = d_zx;
        D.yx[e] = d_yx;
        hx.v[e] = hx.v[e] - P.ch[0] * (d_zy - d_yz);
        hy.v[e] = hy.v[e] - P.ch[1] * (d_xz - d_zx);
        hz.v[e] = hz.v[e] - P.ch[2] * (d_yx - d_xy);
      }
      // CPML slabs, registration order: psi_H in place (only the owner touches it); an x slab corrects Hy and Hz,
      // a y slab Hz and Hx, a z slab Hx and Hy      (fdtd/boundaries.py:461-487, 421-431)
      const int ih = i - 1;
      for (int s = 0; hit_prev != 0 && s < P.n_sl; ++s) {
        if (!((hit_prev >> s) & 1u)) continue;
        const typename FusedParams<T>::Slab& S = P.sl[s];
        // (the stage of plane i carries psi_H of plane i-1; the last iteration of the last chunk has no stage)
        const T* sp = (FDTD_FUSED_PSI_STAGE && s == zs && i < P.x1) ? sH + Lay::FIELD_WORDS + 2 * Lay::PSI_ARR + zs_off : nullptr;
        const i64 idx = S.axis == 0 ? (i64)(ih - S.xs) * plane + p
                                    : (S.axis == 1 ? ((i64)ih * S.t + (j - S.lo)) * Nz + k0
                                                   : ((i64)ih * P.Ny + j) * S.tp + (k0 - S.lo_al));
        const int l0 = S.axis == 0 ? ih - S.lo : (S.axis == 1 ? j - S.lo : k0 - S.lo);
        fused_slab_update<T, VEC, false>(S, S.psiH, S.psiH, true, sp, Lay::PSI_ARR, idx, l0, D, hx, hy, hz, P.ch,
                                           tabs + s * 4 * Lay::TAB_T, early_h && s == sx_prev, xha, xhb);
      }
      const i64 om = off - plane;
      stv<T, VEC>(P.Hout[0] + om, hx);
      stv<T, VEC>(P.Hout[1] + om, hy);
      stv<T, VEC>(P.Hout[2] + om, hz);
    }

    // ---- publish E_new[i] to the block, carry the planes ---------------------------------------------
    if (active) {
      T* x = xch + (i & 1) * Lay::X_WORDS + (r * EV + l) * VEC;
      stv<T, VEC>(x, e0);
      if (!Lay::SHFL) stv<T, VEC>(x + XC, e1);
      stv<T, VEC>(x + XZ, e2);
      ep0 = e0;
      ep1 = e1;
      ep2 = e2;
      if (inside && i < P.x1) {
        hp0 = h0;
        hp1 = h1;
        hp2 = h2;
      }
    }
    hit_prev = hit_now;
    sx_prev = sx;
  }
}

}  // namespace fdtd
