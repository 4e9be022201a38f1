// yee_kernels.cuh -- sm_100a kernels of the fused Yee half-steps.
//
// One kernel per half-step (template IS_E selects E or H).  A thread owns VEC
// consecutive z-cells of one (y) row and marches along x over a chunk of planes,
// carrying the x-neighbour plane in registers, so that every field word crosses
// HBM once per half-step: 128-bit coalesced loads of the own cells, the y-neighbour
// row (an L1 hit inside the block's tile) and one scalar for the z-neighbour.
// Material coefficients, CPML psi updates and field corrections are folded into
// the same pass, selected by a per-(plane, tile) class byte, so homogeneous
// interior tiles stream nothing but the 9 field words of the half-step.
//
// Arithmetic order follows the reference operation by operation (the file is
// compiled with -fmad=false): in float64 the results are bit-identical to the
// reference's numpy backend, in float32 to its true-float32 torch run.
//
// Reference semantics restated here (see oracle/yee_oracle.py for the CPU port):
//   curl_H / curl_E            fdtd/grid.py:29-76
//   E += sc*eps^-1*curl        fdtd/grid.py:283      H -= sc*mu^-1*curl   fdtd/grid.py:309
//   Object.update_E            fdtd/objects.py:118-129
//   AbsorbingObject.update_E   fdtd/objects.py:207-221
//   PML.update_phi_E/H         fdtd/boundaries.py:433-487
//   PML.update_E/H             fdtd/boundaries.py:409-431
#pragma once

// tuning knobs (defaults are the measured best on B200, see profiles/)
#ifndef FDTD_MIN_BLOCKS
#define FDTD_MIN_BLOCKS 3      // resident 256-thread blocks per SM the register budget is sized for
#endif
#ifndef FDTD_MAT_MIN_BLOCKS
#define FDTD_MAT_MIN_BLOCKS 2   // instantiations with material / object code (MAT = true): 122 registers and no spill at two
                                // blocks per SM instead of 80 with spills at three -- 512^3 absorber + lens +0.6 %, the GRIN
                                // slab of config 5 +6.3 % (profiles/r2_mat_variants.txt)
#endif
#ifndef FDTD_BLOCK_THREADS
#define FDTD_BLOCK_THREADS 256 // threads per block of the half-step kernels
#endif
#ifndef FDTD_PREFETCH_PLANES
#define FDTD_PREFETCH_PLANES 1 // software prefetch into L2 this many x-planes ahead (0 = off)
#endif

#ifndef FDTD_PREFETCH_CAP
#define FDTD_PREFETCH_CAP 1    // 1: never prefetch past the block's own x-chunk (the next chunk belongs to a block that
                               // runs much later: by then the lines are evicted again and were fetched for nothing).
                               // 1024^3 f32: DRAM reads 28.36 -> 27.44 GB per launch, 12.91 -> 12.62 ms per step
                               // (profiles/r2_prefetch_cap.txt)
#endif

#ifndef FDTD_PREFETCH_WHAT
#define FDTD_PREFETCH_WHAT 3   // bit 0: the differentiated field G, bit 1: the updated field F
#endif
#ifndef FDTD_STREAM_HINTS
#define FDTD_STREAM_HINTS 0    // 1: evict-first loads/stores (ld.global.cs / st.global.cs) for the updated field
#endif
#ifndef FDTD_H_DOWNWARD
#define FDTD_H_DOWNWARD 0      // 1: the H half-step marches from high x to low x (mirror image of the E half-step)
#endif
#ifndef FDTD_NOINLINE_SLABS
#define FDTD_NOINLINE_SLABS 0  // 1: the CPML pass is an out-of-line call, its registers do not count in the hot loop
#endif
#if defined(FDTD_EMU)
#define FDTD_RARE_FN inline
#elif defined(FDTD_POST_INLINE)
#define FDTD_RARE_FN __device__ __forceinline__
#else
#define FDTD_RARE_FN __device__ __noinline__   // rare paths out of line: their registers do not count in the hot loop
#endif
// absorbing / anisotropic / overlap tiles: inlined (measured +3 % on the 512^3 absorber + lens config over an
// out-of-line call, profiles/r1_special_inline.txt); homogeneous grids never compile this code (MAT = false)
#if defined(FDTD_EMU)
#define FDTD_SPECIAL_FN inline
#elif defined(FDTD_SPECIAL_NOINLINE)
#define FDTD_SPECIAL_FN __device__ __noinline__
#else
#define FDTD_SPECIAL_FN __device__ __forceinline__
#endif
#if defined(FDTD_EMU)
#define FDTD_GRID_CONSTANT
#define FDTD_SLAB_FN inline
#elif FDTD_NOINLINE_SLABS
#define FDTD_GRID_CONSTANT __grid_constant__
#define FDTD_SLAB_FN __device__ __noinline__
#else
#define FDTD_GRID_CONSTANT __grid_constant__
#define FDTD_SLAB_FN __device__ __forceinline__
#endif

namespace fdtd {

typedef long long i64;

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
  T v[VEC];
};

template <typename T, int VEC>
FDTD_DEV Pack<T, VEC> ldv(const T* p) {
  return *reinterpret_cast<const Pack<T, VEC>*>(p);
}
template <typename T, int VEC>
FDTD_DEV void stv(T* p, const Pack<T, VEC>& x) {
  *reinterpret_cast<Pack<T, VEC>*>(p) = x;
}

// streaming (evict-first) access for data touched exactly once per half-step
template <typename T, int VEC>
FDTD_DEV Pack<T, VEC> ldv_stream(const T* p) {
#if !defined(FDTD_EMU) && FDTD_STREAM_HINTS
  if constexpr (sizeof(Pack<T, VEC>) == 16) {
    float4 r = __ldcs(reinterpret_cast<const float4*>(p));
    return *reinterpret_cast<Pack<T, VEC>*>(&r);
  }
#endif
  return ldv<T, VEC>(p);
}
template <typename T, int VEC>
FDTD_DEV void stv_stream(T* p, const Pack<T, VEC>& x) {
#if !defined(FDTD_EMU) && FDTD_STREAM_HINTS
  if constexpr (sizeof(Pack<T, VEC>) == 16) {
    __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&x));
    return;
  }
#endif
  stv<T, VEC>(p, x);
}

// fire-and-forget L2 prefetch of the line holding p (no registers, no scoreboard)
FDTD_DEV void prefetch_l2(const void* p) {
#if !defined(FDTD_EMU) && FDTD_PREFETCH_PLANES > 0
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// T = storage type of the state (fields, psi, detector rings); A = arithmetic type, which is also the type of
// every coefficient (tables, material arrays, source profiles and waveforms).  A = T except in the float32x
// mode (FDTD_F32X): float storage, double arithmetic, one rounding at each store.
template <typename T, typename A = T>
struct SlabK {
  int axis, lo, t, fused, x0, x1;
  int lo_al, tp;  // z slabs: psi rows are [tp] long and start at z = lo_al (both multiples of 4) -> vector access
  i64 count;
  T* psi;      // psi_E in the E half-step, psi_H in the H half-step
  const A* b;  // bE / bH
  const A* c;  // cE / cH
};

// a source folded into the half-step kernel: points (soft, +=) or box (hard, =)
template <typename A>
struct SrcK {
  int kind, comp, n;
  int bb[6];          // local bounding box x0,x1,y0,y1,z0,z1 (half-open); the box itself for FDTD_SRC_BOX
  const i64* idx;     // ascending local linear cell indices
  const A* profile;
  A amplitude;
  const A* wave;
  i64 w;              // waveform table index (+ dyn[0])
};

// a detector folded into the half-step kernel
template <typename T>
struct DetK {
  int n;
  int bb[6];
  const i64* idx;     // ascending
  const int* pos;     // ring position of each sorted entry
  T* ring;
  i64 slot;           // (+ dyn[1])
};

template <typename T, typename A = T>
struct HalfStepParams {
  int Nx, Ny, Nz, x_offset, Nx_global;
  int x_begin, x_end, x_chunk;
  int lanes_z, lanes_shift, rows;  // block = lanes_z * rows threads, thread -> (row, lane)
  i64 plane;
  int y_begin, y_end, z_begin, z_end;  // cell box of this launch in y and z (full grid unless a shell launch)
  T* F[3];        // field being updated (read)
  T* Fo[3];       // where the updated field is written: F itself, or the other buffer of a ping-pong pair
  const T* G[3];  // field being differentiated
  A sc;
  A bg_c[3];      // sc * background inverse material, rounded as the reference rounds it
  A bg_inv[3];    // background inverse material itself (AnisotropicObject cells round sc*(inv*curl))
  const A* inv[3];       // effective inverse material of the curl term, or null
  const A* inv2[3];      // second object covering a cell (overlaps), or null
  const A* inv_grid[3];  // grid's own eps^-1 for the PML correction, or null (= inv)
  const A* absorb[3];    // absorption factor, or null
  const A* absorb2[3];   // absorption factor of the second object covering a cell, or null
  const unsigned char* cls;
  unsigned char cls_vary;  // FDTD_CLS_VARY_E or FDTD_CLS_VARY_H
  int n_slabs;
  SlabK<T, A> slabs[6];
  // sources and detectors folded into this pass (only when nothing has to run between the field
  // update and them, i.e. no periodic copy / late PML correction); registration order
  // boundary plane pushed straight into the neighbour slab's ghost plane over NVLink (peer pointers
  // from CUDA IPC): components y and z of plane `push_plane`, written by the same threads that compute them
  int push_plane;
  T* push_y;
  T* push_z;
  int n_src, n_det;
  const i64* dyn;  // optional device int64[2] {waveform index base, ring slot base} (CUDA-graph replays)
  SrcK<A> src[FDTD_FUSED_MAX];
  DetK<T> det[FDTD_FUSED_MAX];
};

// load VEC stored values and widen them to the arithmetic type / narrow and store (the one rounding per store)
template <typename T, typename A, int VEC>
FDTD_DEV void ld_wide(const T* p, A (&out)[VEC]) {
  const Pack<T, VEC> v = ldv<T, VEC>(p);
#pragma unroll
  for (int e = 0; e < VEC; ++e) out[e] = (A)v.v[e];
}
template <typename T, typename A, int VEC>
FDTD_DEV void st_narrow(T* p, const A (&in)[VEC]) {
  Pack<T, VEC> v;
#pragma unroll
  for (int e = 0; e < VEC; ++e) v.v[e] = (T)in[e];
  stv<T, VEC>(p, v);
}

// CPML update of one slab for the VEC cells of a thread.  A = slab axis; (A, U, W) cyclic.
// d0 = one-sided difference of G_W along A (drives psi[0], corrects F_U with sign -),
// d1 = one-sided difference of G_U along A (drives psi[1], corrects F_W with sign +).
template <typename T, typename A, int VEC, bool IS_E>
FDTD_DEV void slab_cells(const SlabK<T, A>& S, i64 idx0, int l0, int lstep, const A (&d0)[VEC],
                         const A (&d1)[VEC], A (&fu)[VEC], A (&fw)[VEC], const A (&cu)[VEC],
                         const A (&cw)[VEC]) {
  // x / y slabs: the VEC cells share l (lstep = 0).  z slabs: l = l0 + e; psi rows are padded so that
  // idx0 is vector-aligned, cells outside [0, t) are padding (stay zero).  Always 128-bit psi access.
  Pack<T, VEC> a = ldv<T, VEC>(S.psi + idx0);
  Pack<T, VEC> b = ldv<T, VEC>(S.psi + S.count + idx0);
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const int l = l0 + e * lstep;
    if (l >= 0 && l < S.t) {
      const A bb = S.b[l];
      const A cc = S.c[l];
      // psi *= b; psi[inner] += diff * c[inner]      fdtd/boundaries.py:439-454, 467-482
      const bool inner = IS_E ? (l >= 1) : (l < S.t - 1);
      A p0 = (A)a.v[e] * bb;
      A p1 = (A)b.v[e] * bb;
      if (inner) {
        p0 = p0 + d0[e] * cc;
        p1 = p1 + d1[e] * cc;
      }
      a.v[e] = (T)p0;
      b.v[e] = (T)p1;
      if (S.fused) {
        // phi_U = 0 - psi0, phi_W = psi1 - 0; F[loc] +-= sc*inv*phi   fdtd/boundaries.py:409-431, 456-459
        // (float32x: the correction uses the psi just computed, before it is rounded for storage)
        const A phi_u = A(0) - p0;
        const A phi_w = p1 - A(0);
        if (IS_E) {
          fu[e] = fu[e] + cu[e] * phi_u;
          fw[e] = fw[e] + cw[e] * phi_w;
        } else {
          fu[e] = fu[e] - cu[e] * phi_u;
          fw[e] = fw[e] - cw[e] * phi_w;
        }
      }
    }
  }
  stv<T, VEC>(S.psi + idx0, a);
  stv<T, VEC>(S.psi + S.count + idx0, b);
}

FDTD_DEV int lower_bound_i64(const i64* a, int n, i64 key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// sources (registration order) then detectors on the final values of a thread's cells
// (fdtd/grid.py:294-299, 320-325; fdtd/sources.py:93-109, 278-297, 476-486; fdtd/detectors.py:114-124)
template <typename A, int VEC>
struct CellFields {
  A fx[VEC], fy[VEC], fz[VEC];
};

template <typename T, typename A, int VEC>
FDTD_RARE_FN void fused_post(const HalfStepParams<T, A>& P, int i, int j, int k0, i64 lin0, CellFields<A, VEC>& V) {
  A (&fx)[VEC] = V.fx;
  A (&fy)[VEC] = V.fy;
  A (&fz)[VEC] = V.fz;
  for (int s = 0; s < P.n_src; ++s) {
    const SrcK<A>& S = P.src[s];
    if (i < S.bb[0] || i >= S.bb[1] || j < S.bb[2] || j >= S.bb[3] || k0 + VEC <= S.bb[4] || k0 >= S.bb[5]) continue;
    const A wv = S.wave[S.w + (P.dyn ? P.dyn[0] : 0)];
    if (S.kind == FDTD_SRC_BOX) {
      // (a hard source writes a STORED value: rounded to the storage type like the field it replaces)
      const A v = (A)(T)(S.amplitude * wv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (k0 + e >= S.bb[4] && k0 + e < S.bb[5]) {
          if (S.comp == 0) fx[e] = v; else if (S.comp == 1) fy[e] = v; else fz[e] = v;
        }
      }
    } else {
      for (int n = lower_bound_i64(S.idx, S.n, lin0); n < S.n && S.idx[n] < lin0 + VEC; ++n) {
        const int de = (int)(S.idx[n] - lin0);
        const A v = S.profile[n] * wv;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (e == de) {
            if (S.comp == 0) fx[e] = fx[e] + v; else if (S.comp == 1) fy[e] = fy[e] + v; else fz[e] = fz[e] + v;
          }
        }
      }
    }
  }
  for (int s = 0; s < P.n_det; ++s) {
    const DetK<T>& D = P.det[s];
    if (i < D.bb[0] || i >= D.bb[1] || j < D.bb[2] || j >= D.bb[3] || k0 + VEC <= D.bb[4] || k0 >= D.bb[5]) continue;
    const i64 slot = D.slot + (P.dyn ? P.dyn[1] : 0);
    for (int n = lower_bound_i64(D.idx, D.n, lin0); n < D.n && D.idx[n] < lin0 + VEC; ++n) {
      const int de = (int)(D.idx[n] - lin0);
      T* out = D.ring + (slot * D.n + D.pos[n]) * 3;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (e == de) {
          out[0] = fx[e];
          out[1] = fy[e];
          out[2] = fz[e];
        }
      }
    }
  }
}

// what the CPML pass needs about the VEC cells of a thread
template <typename A, int VEC>
struct CellState {
  A d_zy[VEC], d_yz[VEC], d_xz[VEC], d_zx[VEC], d_yx[VEC], d_xy[VEC];
  A fx[VEC], fy[VEC], fz[VEC];
  A cx[VEC], cy[VEC], cz[VEC];
};

// E update of the cells of a tile that an AbsorbingObject, an AnisotropicObject or overlapping objects touch
// (fdtd/objects.py:118-129, 207-221, 254-269).  C.c* = sc * effective eps^-1 as computed by the caller.
//
// The reference updates every object in registration order; a cell covered by two objects gets two updates.
// The host bakes the first object covering a cell into layer 1 (inv / absorb) and the second into layer 2
// (inv2 / absorb2); each layer is applied exactly as its object kind does it:
//   plain        E += (sc*eps^-1) * curl
//   anisotropic  E += sc * (eps^-1 * curl)         the product with curl is rounded BEFORE the scaling by sc;
//                                                   marked by a NEGATIVE zero in the grid's eps^-1: x-component
//                                                   for layer 1, y-component for layer 2
//   absorbing    E *= (1-f)/(1+f); E += (sc*eps^-1)*curl / (1+f)      (f = 0 gives the plain update bit for bit)
template <typename T, typename A, int VEC>
FDTD_SPECIAL_FN void special_update(const HalfStepParams<T, A>& P, CellState<A, VEC>& C, i64 off, unsigned cls) {
  A tx[VEC], ty[VEC], tz[VEC], ux[VEC], uy[VEC], uz[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    ux[e] = C.d_zy[e] - C.d_yz[e];
    uy[e] = C.d_xz[e] - C.d_zx[e];
    uz[e] = C.d_yx[e] - C.d_xy[e];
    tx[e] = C.cx[e] * ux[e];
    ty[e] = C.cy[e] * uy[e];
    tz[e] = C.cz[e] * uz[e];
  }
  if (cls & FDTD_CLS_ANISO) {
    const Pack<A, VEC> mark = ldv<A, VEC>(P.inv_grid[0] + off);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if ((mark.v[e] == A(0)) && fdtd_signbit(mark.v[e])) {
        const bool vary = (cls & P.cls_vary) != 0;
        tx[e] = P.sc * ((vary ? P.inv[0][off + e] : P.bg_inv[0]) * ux[e]);
        ty[e] = P.sc * ((vary ? P.inv[1][off + e] : P.bg_inv[1]) * uy[e]);
        tz[e] = P.sc * ((vary ? P.inv[2][off + e] : P.bg_inv[2]) * uz[e]);
      }
    }
  }
  // ---- layer 1
  if (cls & FDTD_CLS_ABSORB) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const A q0 = P.absorb[0][off + e], q1 = P.absorb[1][off + e], q2 = P.absorb[2][off + e];
      C.fx[e] = C.fx[e] * ((A(1) - q0) / (A(1) + q0));
      C.fy[e] = C.fy[e] * ((A(1) - q1) / (A(1) + q1));
      C.fz[e] = C.fz[e] * ((A(1) - q2) / (A(1) + q2));
      C.fx[e] = C.fx[e] + tx[e] / (A(1) + q0);
      C.fy[e] = C.fy[e] + ty[e] / (A(1) + q1);
      C.fz[e] = C.fz[e] + tz[e] / (A(1) + q2);
    }
  } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      C.fx[e] = C.fx[e] + tx[e];
      C.fy[e] = C.fy[e] + ty[e];
      C.fz[e] = C.fz[e] + tz[e];
    }
  }
  // ---- layer 2
  if (cls & FDTD_CLS_OVERLAP) {
    bool aniso2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) aniso2[e] = false;
    if (cls & FDTD_CLS_ANISO) {
      const Pack<A, VEC> mark2 = ldv<A, VEC>(P.inv_grid[1] + off);
#pragma unroll
      for (int e = 0; e < VEC; ++e) aniso2[e] = (mark2.v[e] == A(0)) && fdtd_signbit(mark2.v[e]);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const A b0 = P.inv2[0][off + e], b1 = P.inv2[1][off + e], b2 = P.inv2[2][off + e];
      tx[e] = aniso2[e] ? P.sc * (b0 * ux[e]) : (P.sc * b0) * ux[e];
      ty[e] = aniso2[e] ? P.sc * (b1 * uy[e]) : (P.sc * b1) * uy[e];
      tz[e] = aniso2[e] ? P.sc * (b2 * uz[e]) : (P.sc * b2) * uz[e];
    }
    if (cls & FDTD_CLS_ABSORB2) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const A q0 = P.absorb2[0][off + e], q1 = P.absorb2[1][off + e], q2 = P.absorb2[2][off + e];
        C.fx[e] = C.fx[e] * ((A(1) - q0) / (A(1) + q0));
        C.fy[e] = C.fy[e] * ((A(1) - q1) / (A(1) + q1));
        C.fz[e] = C.fz[e] * ((A(1) - q2) / (A(1) + q2));
        C.fx[e] = C.fx[e] + tx[e] / (A(1) + q0);
        C.fy[e] = C.fy[e] + ty[e] / (A(1) + q1);
        C.fz[e] = C.fz[e] + tz[e] / (A(1) + q2);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        C.fx[e] = C.fx[e] + tx[e];
        C.fy[e] = C.fy[e] + ty[e];
        C.fz[e] = C.fz[e] + tz[e];
      }
    }
  }
}

// all CPML slabs a thread's cells belong to, in registration order
template <typename T, typename A, int VEC, bool IS_E>
FDTD_SLAB_FN void slab_pass(const HalfStepParams<T, A>& P, CellState<A, VEC>& C, int i, int j, int k0, i64 p,
                            i64 off, unsigned cls) {
  const int ig = i + P.x_offset;
  if (IS_E && (cls & FDTD_CLS_OBJECT) && P.inv_grid[0] != nullptr) {
    // the correction uses the GRID's eps^-1, which is zero inside objects (fdtd/objects.py:92)
    Pack<A, VEC> a0 = ldv<A, VEC>(P.inv_grid[0] + off);
    Pack<A, VEC> a1 = ldv<A, VEC>(P.inv_grid[1] + off);
    Pack<A, VEC> a2 = ldv<A, VEC>(P.inv_grid[2] + off);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      C.cx[e] = P.sc * a0.v[e];
      C.cy[e] = P.sc * a1.v[e];
      C.cz[e] = P.sc * a2.v[e];
    }
  }
  for (int s = 0; s < P.n_slabs; ++s) {
    const SlabK<T, A>& S = P.slabs[s];
    if (S.axis == 0) {
      if (i >= S.x0 && i < S.x1)
        slab_cells<T, A, VEC, IS_E>(S, (i64)(i - S.x0) * P.plane + p, ig - S.lo, 0, C.d_zx, C.d_yx, C.fy, C.fz,
                                 C.cy, C.cz);
    } else if (S.axis == 1) {
      const int l = j - S.lo;
      if (l >= 0 && l < S.t)
        slab_cells<T, A, VEC, IS_E>(S, ((i64)i * S.t + l) * P.Nz + k0, l, 0, C.d_xy, C.d_zy, C.fz, C.fx, C.cz,
                                 C.cx);
    } else {
      const int l0 = k0 - S.lo;
      if (l0 + VEC > 0 && l0 < S.t)
        slab_cells<T, A, VEC, IS_E>(S, ((i64)i * P.Ny + j) * S.tp + (k0 - S.lo_al), l0, 1, C.d_yz, C.d_xz, C.fx,
                                 C.fy, C.cx, C.cy);
    }
  }
}

// HAS_POST: sources / detectors folded into this pass (small grids, where a separate launch per source
// and detector would dominate the step); compiled out otherwise
// HAS_PUSH: the launch covers the slab's boundary plane and stores it into the neighbour's ghost plane as well
// (compute + halo transfer in one kernel; fdtd_halo_signal publishes it afterwards)
// MAT: the grid has material arrays and a tile-class map; without them (homogeneous grids, e.g. the 1024^3
// benchmark) all coefficient / object / absorber code is compiled out and costs no registers
template <typename T, int VEC, bool IS_E, bool HAS_POST, bool HAS_PUSH, bool MAT, typename A = T>
__global__ void __launch_bounds__(FDTD_BLOCK_THREADS, (sizeof(A) > sizeof(T) ? 2 : (MAT ? FDTD_MAT_MIN_BLOCKS : FDTD_MIN_BLOCKS))) halfstep_kernel(const FDTD_GRID_CONSTANT HalfStepParams<T, A> P) {
  const int tid = threadIdx.x;
  const int lane = tid & (P.lanes_z - 1);
  const int row = tid >> P.lanes_shift;
  const int k0 = P.z_begin + (blockIdx.x * P.lanes_z + lane) * VEC;
  const int j = P.y_begin + blockIdx.y * P.rows + row;
  if (j >= P.y_end || k0 >= P.z_end) return;
  const int Nz = P.Nz;
  const i64 plane = P.plane;
  const i64 p = (i64)j * Nz + k0;
  const int i0 = P.x_begin + blockIdx.z * P.x_chunk;
  const int i1 = (i0 + P.x_chunk < P.x_end) ? i0 + P.x_chunk : P.x_end;

  const T* __restrict__ Gx = P.G[0];
  const T* __restrict__ Gy = P.G[1];
  const T* __restrict__ Gz = P.G[2];
  T* Fx = P.F[0];
  T* Fy = P.F[1];
  T* Fz = P.F[2];

  // neighbour offsets: backward differences for E (fdtd/grid.py:66-74), forward for H (:41-49)
  const i64 off_y = IS_E ? -(i64)Nz : (i64)Nz;
  const i64 off_zs = IS_E ? -1 : VEC;
  const bool my = IS_E ? (j >= 1) : (j < P.Ny - 1);
  bool mz[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) mz[e] = IS_E ? (k0 + e >= 1) : (k0 + e < Nz - 1);

  // loop-invariant slab membership in y and z
  bool pml_yz = false;
  for (int s = 0; s < P.n_slabs; ++s) {
    const SlabK<T, A>& S = P.slabs[s];
    if (S.axis == 1) pml_yz |= (j >= S.lo) && (j < S.lo + S.t);
    if (S.axis == 2) pml_yz |= (k0 + VEC > S.lo) && (k0 < S.lo + S.t);
  }

  // loop-invariant: does any folded source / detector touch this thread's row and z-range?
  bool post_yz = false;
  for (int s = 0; HAS_POST && s < P.n_src; ++s)
    post_yz |= (j >= P.src[s].bb[2]) && (j < P.src[s].bb[3]) && (k0 + VEC > P.src[s].bb[4]) && (k0 < P.src[s].bb[5]);
  for (int s = 0; HAS_POST && s < P.n_det; ++s)
    post_yz |= (j >= P.det[s].bb[2]) && (j < P.det[s].bb[3]) && (k0 + VEC > P.det[s].bb[4]) && (k0 < P.det[s].bb[5]);

  // x-neighbour plane carried in registers.  OWN_LOAD (E; H when marching downward): every plane's own
  // (Gy, Gz) are loaded and the neighbour plane (i-1 for E, i+1 for H) is last iteration's own.  Otherwise
  // (H marching upward): plane i comes from the carry and plane i+1 is loaded.
  constexpr bool DOWN = !IS_E && (FDTD_H_DOWNWARD != 0);
  constexpr bool OWN_LOAD = IS_E || DOWN;
  constexpr int PF_DIR = DOWN ? -1 : 1;
  Pack<T, VEC> carry_y, carry_z;
  {
    const i64 o = (IS_E ? (i64)(i0 - 1) : (DOWN ? (i64)i1 : (i64)i0)) * plane + p;
    carry_y = ldv<T, VEC>(Gy + o);
    carry_z = ldv<T, VEC>(Gz + o);
  }
  const i64 cls_stride = (i64)gridDim.y * gridDim.x;
  const i64 cls_tile = (i64)blockIdx.y * gridDim.x + blockIdx.x;

  for (int it = 0; it < i1 - i0; ++it) {
    const int i = DOWN ? i1 - 1 - it : i0 + it;
    const i64 off = (i64)i * plane + p;
    const unsigned cls = (MAT && P.cls) ? P.cls[(i64)i * cls_stride + cls_tile] : 0u;

#if FDTD_PREFETCH_PLANES > 0
    if (FDTD_PREFETCH_CAP ? (DOWN ? (i - FDTD_PREFETCH_PLANES >= i0) : (i + FDTD_PREFETCH_PLANES < i1))
                          : (DOWN ? (i - FDTD_PREFETCH_PLANES >= 0) : (i + FDTD_PREFETCH_PLANES < P.Nx))) {
      const i64 pf = off + (i64)(PF_DIR * FDTD_PREFETCH_PLANES) * plane;
      if (FDTD_PREFETCH_WHAT & 1) {
        prefetch_l2(Gx + pf);
        prefetch_l2(Gy + pf + (OWN_LOAD ? 0 : plane));
        prefetch_l2(Gz + pf + (OWN_LOAD ? 0 : plane));
      }
      if (FDTD_PREFETCH_WHAT & 2) {
        prefetch_l2(Fx + pf);
        prefetch_l2(Fy + pf);
        prefetch_l2(Fz + pf);
      }
    }
#endif
    // ---- loads -------------------------------------------------------------------------
    Pack<T, VEC> gx = ldv<T, VEC>(Gx + off);
    Pack<T, VEC> gy, gz, xnb_y, xnb_z;
    if (OWN_LOAD) {
      gy = ldv<T, VEC>(Gy + off);
      gz = ldv<T, VEC>(Gz + off);
      xnb_y = carry_y;
      xnb_z = carry_z;
    } else {
      gy = carry_y;
      gz = carry_z;
      xnb_y = ldv<T, VEC>(Gy + off + plane);
      xnb_z = ldv<T, VEC>(Gz + off + plane);
    }
    const Pack<T, VEC> ynb_x = ldv<T, VEC>(Gx + off + off_y);
    const Pack<T, VEC> ynb_z = ldv<T, VEC>(Gz + off + off_y);
    const T zs_x = Gx[off + off_zs];
    const T zs_y = Gy[off + off_zs];
    Pack<T, VEC> f0 = ldv_stream<T, VEC>(Fx + off);
    Pack<T, VEC> f1 = ldv_stream<T, VEC>(Fy + off);
    Pack<T, VEC> f2 = ldv_stream<T, VEC>(Fz + off);

    const int ig = i + P.x_offset;
    const bool mx = IS_E ? (ig >= 1) : (ig < P.Nx_global - 1);

    // ---- one-sided differences d_ca = d G_c / d a, each masked independently ------------
    // (float32x: the stored values are widened here, where they are used; the loaded vectors stay narrow)
    A d_zy[VEC], d_yz[VEC], d_xz[VEC], d_zx[VEC], d_yx[VEC], d_xy[VEC];
    A fx[VEC], fy[VEC], fz[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const A znb_x = (A)(IS_E ? (e == 0 ? zs_x : gx.v[e > 0 ? e - 1 : 0])
                               : (e == VEC - 1 ? zs_x : gx.v[e < VEC - 1 ? e + 1 : 0]));
      const A znb_y = (A)(IS_E ? (e == 0 ? zs_y : gy.v[e > 0 ? e - 1 : 0])
                               : (e == VEC - 1 ? zs_y : gy.v[e < VEC - 1 ? e + 1 : 0]));
      const A ax = (A)gx.v[e], ay = (A)gy.v[e], az = (A)gz.v[e];
      if (IS_E) {
        d_zy[e] = my ? az - (A)ynb_z.v[e] : A(0);
        d_xy[e] = my ? ax - (A)ynb_x.v[e] : A(0);
        d_yz[e] = mz[e] ? ay - znb_y : A(0);
        d_xz[e] = mz[e] ? ax - znb_x : A(0);
        d_zx[e] = mx ? az - (A)xnb_z.v[e] : A(0);
        d_yx[e] = mx ? ay - (A)xnb_y.v[e] : A(0);
      } else {
        d_zy[e] = my ? (A)ynb_z.v[e] - az : A(0);
        d_xy[e] = my ? (A)ynb_x.v[e] - ax : A(0);
        d_yz[e] = mz[e] ? znb_y - ay : A(0);
        d_xz[e] = mz[e] ? znb_x - ax : A(0);
        d_zx[e] = mx ? (A)xnb_z.v[e] - az : A(0);
        d_yx[e] = mx ? (A)xnb_y.v[e] - ay : A(0);
      }
      fx[e] = (A)f0.v[e];
      fy[e] = (A)f1.v[e];
      fz[e] = (A)f2.v[e];
    }

    // ---- coefficients --------------------------------------------------------------------
    A cx[VEC], cy[VEC], cz[VEC];
    if (cls & P.cls_vary) {
      // material arrays have no ghost planes: [x][y][z] index == off
      Pack<A, VEC> a0 = ldv<A, VEC>(P.inv[0] + off);
      Pack<A, VEC> a1 = ldv<A, VEC>(P.inv[1] + off);
      Pack<A, VEC> a2 = ldv<A, VEC>(P.inv[2] + off);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        cx[e] = P.sc * a0.v[e];
        cy[e] = P.sc * a1.v[e];
        cz[e] = P.sc * a2.v[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        cx[e] = P.bg_c[0];
        cy[e] = P.bg_c[1];
        cz[e] = P.bg_c[2];
      }
    }

    // ---- field update ----------------------------------------------------------------------
    if (IS_E && (cls & (FDTD_CLS_ABSORB | FDTD_CLS_ANISO | FDTD_CLS_OVERLAP | FDTD_CLS_ABSORB2))) {
      // tiles an AbsorbingObject, an AnisotropicObject or two overlapping objects touch: out-of-line, so that
      // their registers and divisions do not burden the streaming path
      CellState<A, VEC> C;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        C.d_zy[e] = d_zy[e]; C.d_yz[e] = d_yz[e]; C.d_xz[e] = d_xz[e];
        C.d_zx[e] = d_zx[e]; C.d_yx[e] = d_yx[e]; C.d_xy[e] = d_xy[e];
        C.fx[e] = fx[e]; C.fy[e] = fy[e]; C.fz[e] = fz[e];
        C.cx[e] = cx[e]; C.cy[e] = cy[e]; C.cz[e] = cz[e];
      }
      special_update<T, A, VEC>(P, C, off, cls);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        fx[e] = C.fx[e]; fy[e] = C.fy[e]; fz[e] = C.fz[e];
      }
    } else {
      // F +-= (sc * inverse material) * curl      fdtd/grid.py:283, 309; fdtd/objects.py:118-129
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const A tx = cx[e] * (d_zy[e] - d_yz[e]);
        const A ty = cy[e] * (d_xz[e] - d_zx[e]);
        const A tz = cz[e] * (d_yx[e] - d_xy[e]);
        if (IS_E) {
          fx[e] = fx[e] + tx;
          fy[e] = fy[e] + ty;
          fz[e] = fz[e] + tz;
        } else {
          fx[e] = fx[e] - tx;
          fy[e] = fy[e] - ty;
          fz[e] = fz[e] - tz;
        }
      }
    }

    // ---- CPML slabs (registration order) -----------------------------------------------------
    bool pml_x = false;
    for (int s = 0; s < P.n_slabs; ++s) {
      const SlabK<T, A>& S = P.slabs[s];
      if (S.axis == 0) pml_x |= (i >= S.x0) && (i < S.x1);
    }
    if (pml_x || pml_yz) {
      CellState<A, VEC> C;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        C.d_zy[e] = d_zy[e]; C.d_yz[e] = d_yz[e]; C.d_xz[e] = d_xz[e];
        C.d_zx[e] = d_zx[e]; C.d_yx[e] = d_yx[e]; C.d_xy[e] = d_xy[e];
        C.fx[e] = fx[e]; C.fy[e] = fy[e]; C.fz[e] = fz[e];
        C.cx[e] = cx[e]; C.cy[e] = cy[e]; C.cz[e] = cz[e];
      }
      slab_pass<T, A, VEC, IS_E>(P, C, i, j, k0, p, off, cls);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        fx[e] = C.fx[e]; fy[e] = C.fy[e]; fz[e] = C.fz[e];
      }
    }

    // ---- folded sources and detectors (rare: only threads whose row / z-range a source or detector touches)
    if (HAS_POST && post_yz) {
      CellFields<A, VEC> V;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        V.fx[e] = fx[e]; V.fy[e] = fy[e]; V.fz[e] = fz[e];
      }
      fused_post<T, A, VEC>(P, i, j, k0, off, V);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        fx[e] = V.fx[e]; fy[e] = V.fy[e]; fz[e] = V.fz[e];
      }
    }

    // ---- stores --------------------------------------------------------------------------------
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      f0.v[e] = (T)fx[e];
      f1.v[e] = (T)fy[e];
      f2.v[e] = (T)fz[e];
    }
    stv_stream<T, VEC>(P.Fo[0] + off, f0);
    stv_stream<T, VEC>(P.Fo[1] + off, f1);
    stv_stream<T, VEC>(P.Fo[2] + off, f2);
    if (HAS_PUSH && i == P.push_plane) {
      stv<T, VEC>(P.push_y + p, f1);
      stv<T, VEC>(P.push_z + p, f2);
    }

    if (OWN_LOAD) {
      carry_y = gy;
      carry_z = gz;
    } else {
      carry_y = xnb_y;
      carry_z = xnb_z;
    }
  }
}

// ---- post kernels ------------------------------------------------------------------------------

// periodic copy of one plane of all three components (fdtd/boundaries.py:184-219)
template <typename T>
__global__ void periodic_kernel(T* F0, T* F1, T* F2, int axis, int Nx, int Ny, int Nz, i64 plane,
                                int src, int dst) {
  const i64 n_a = axis == 0 ? Ny : Nx;
  const i64 n_b = axis == 2 ? Ny : Nz;
  const i64 total = n_a * n_b * 3;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (i64)gridDim.x * blockDim.x) {
    const int c = (int)(t / (n_a * n_b));
    const i64 r = t % (n_a * n_b);
    const i64 a = r / n_b, b = r % n_b;
    i64 so, dofs;
    if (axis == 0) {
      so = (i64)src * plane + a * Nz + b;
      dofs = (i64)dst * plane + a * Nz + b;
    } else if (axis == 1) {
      so = a * plane + (i64)src * Nz + b;
      dofs = a * plane + (i64)dst * Nz + b;
    } else {
      so = a * plane + b * Nz + src;
      dofs = a * plane + b * Nz + dst;
    }
    T* F = c == 0 ? F0 : (c == 1 ? F1 : F2);
    F[dofs] = F[so];
  }
}

// field correction of a PML registered after a periodic boundary (fdtd/boundaries.py:409-431)
template <typename T, bool IS_E, typename A = T>
__global__ void pml_add_kernel(SlabK<T, A> S, T* F0, T* F1, T* F2, const A* c0, const A* c1,
                               const A* c2, A bg0, A bg1, A bg2, A sc, int Nx, int Ny, int Nz,
                               i64 plane) {
  for (i64 n = (i64)blockIdx.x * blockDim.x + threadIdx.x; n < S.count;
       n += (i64)gridDim.x * blockDim.x) {
    int i, j, k;
    if (S.axis == 0) {
      i = S.x0 + (int)(n / plane);
      j = (int)((n % plane) / Nz);
      k = (int)(n % Nz);
    } else if (S.axis == 1) {
      i = (int)(n / ((i64)S.t * Nz));
      j = S.lo + (int)((n / Nz) % S.t);
      k = (int)(n % Nz);
    } else {
      i = (int)(n / ((i64)Ny * S.tp));
      j = (int)((n / S.tp) % Ny);
      k = S.lo_al + (int)(n % S.tp);
      if (k < S.lo || k >= S.lo + S.t) continue;  // row padding
    }
    const i64 off = (i64)i * plane + (i64)j * Nz + k;
    const int u = (S.axis + 1) % 3, w = (S.axis + 2) % 3;
    T* Fu = u == 0 ? F0 : (u == 1 ? F1 : F2);
    T* Fw = w == 0 ? F0 : (w == 1 ? F1 : F2);
    const A* cu_p = u == 0 ? c0 : (u == 1 ? c1 : c2);
    const A* cw_p = w == 0 ? c0 : (w == 1 ? c1 : c2);
    const A cu = cu_p ? sc * cu_p[off] : (u == 0 ? bg0 : (u == 1 ? bg1 : bg2));
    const A cw = cw_p ? sc * cw_p[off] : (w == 0 ? bg0 : (w == 1 ? bg1 : bg2));
    const A phi_u = A(0) - (A)S.psi[n];
    const A phi_w = (A)S.psi[S.count + n] - A(0);
    if (IS_E) {
      Fu[off] = (T)((A)Fu[off] + cu * phi_u);
      Fw[off] = (T)((A)Fw[off] + cw * phi_w);
    } else {
      Fu[off] = (T)((A)Fu[off] - cu * phi_u);
      Fw[off] = (T)((A)Fw[off] - cw * phi_w);
    }
  }
}

// An object that is the third or later one covering a cell: its update_E on the cells of its box named by `mask`
// (fdtd/objects.py:118-129 plain, 254-269 anisotropic, 207-221 absorbing), with curl_H recomputed from H
// (fdtd/grid.py:54-76: backward differences, each masked on its own face).  H pointers are at local plane 0 of
// arrays with ghost planes, so plane -1 is readable on slabs with a left neighbour.
template <typename T, typename A = T>
__global__ void object_layer_kernel(T* E0, T* E1, T* E2, const T* H0, const T* H1, const T* H2, const A* i0,
                                    const A* i1, const A* i2, const A* a0, const A* a1, const A* a2,
                                    const unsigned char* mask, int kind, int x0, int x1, int y0, int y1, int z0, int z1,
                                    int Nz, i64 plane, int x_offset, A sc) {
  const i64 ny = y1 - y0, nz = z1 - z0;
  const i64 total = (i64)(x1 - x0) * ny * nz;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    if (!mask[t]) continue;
    const int i = x0 + (int)(t / (ny * nz)), j = y0 + (int)((t / nz) % ny), k = z0 + (int)(t % nz);
    const i64 o = (i64)i * plane + (i64)j * Nz + k;
    const bool mx = i + x_offset >= 1, my = j >= 1, mz = k >= 1;
    const A hx = (A)H0[o], hy = (A)H1[o], hz = (A)H2[o];
    const A d_zy = my ? hz - (A)H2[o - Nz] : A(0);
    const A d_xy = my ? hx - (A)H0[o - Nz] : A(0);
    const A d_yz = mz ? hy - (A)H1[o - 1] : A(0);
    const A d_xz = mz ? hx - (A)H0[o - 1] : A(0);
    const A d_zx = mx ? hz - (A)H2[o - plane] : A(0);
    const A d_yx = mx ? hy - (A)H1[o - plane] : A(0);
    const A c0 = d_zy - d_yz, c1 = d_xz - d_zx, c2 = d_yx - d_xy;
    A e0 = (A)E0[o], e1 = (A)E1[o], e2 = (A)E2[o];
    if (kind == FDTD_OBJ_ANISO) {
      e0 = e0 + sc * (i0[t] * c0);
      e1 = e1 + sc * (i1[t] * c1);
      e2 = e2 + sc * (i2[t] * c2);
    } else if (kind == FDTD_OBJ_ABSORB) {
      const A q0 = a0[t], q1 = a1[t], q2 = a2[t];
      e0 = e0 * ((A(1) - q0) / (A(1) + q0));
      e1 = e1 * ((A(1) - q1) / (A(1) + q1));
      e2 = e2 * ((A(1) - q2) / (A(1) + q2));
      e0 = e0 + ((sc * i0[t]) * c0) / (A(1) + q0);
      e1 = e1 + ((sc * i1[t]) * c1) / (A(1) + q1);
      e2 = e2 + ((sc * i2[t]) * c2) / (A(1) + q2);
    } else {
      e0 = e0 + (sc * i0[t]) * c0;
      e1 = e1 + (sc * i1[t]) * c1;
      e2 = e2 + (sc * i2[t]) * c2;
    }
    E0[o] = (T)e0;
    E1[o] = (T)e1;
    E2[o] = (T)e2;
  }
}

// soft source: F[idx[n]] += profile[n] * wave   (fdtd/sources.py:93-109, 278-297)
// (a point list never names a cell twice: the points of a LineSource differ along its longest axis, so the
// read-modify-write needs no atomic)
template <typename T, typename A = T>
__global__ void source_points_kernel(T* F, const i64* idx, const A* profile, int n, const A* wave,
                                     i64 w) {
  const A s = wave[w];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    F[idx[t]] = (T)((A)F[idx[t]] + profile[t] * s);
  }
}

// hard source: F[box] = amplitude * wave   (fdtd/sources.py:476-486)
template <typename T, typename A = T>
__global__ void source_box_kernel(T* F, int x0, int x1, int y0, int y1, int z0, int z1, int Nz,
                                  i64 plane, A amplitude, const A* wave, i64 w) {
  const T v = (T)(amplitude * wave[w]);
  const i64 ny = y1 - y0, nz = z1 - z0;
  const i64 total = (i64)(x1 - x0) * ny * nz;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (i64)gridDim.x * blockDim.x) {
    const i64 x = x0 + t / (ny * nz);
    const i64 y = y0 + (t / nz) % ny;
    const i64 z = z0 + t % nz;
    F[x * plane + y * Nz + z] = v;
  }
}

// detector sampling into the device ring (fdtd/detectors.py:114-124, 241-263)
template <typename T>
__global__ void detector_kernel(const T* F0, const T* F1, const T* F2, const i64* idx, const int* pos, int n,
                                T* ring, i64 slot) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 3 * n; t += gridDim.x * blockDim.x) {
    const int pt = t / 3, c = t % 3;
    const T* F = c == 0 ? F0 : (c == 1 ? F1 : F2);
    ring[(slot * n + pos[pt]) * 3 + c] = F[idx[pt]];
  }
}

// running DFT of a filled ring part: acc[f][v] += sum_s ring[s][v] * tw[s][f]  (complex, float64, rows in order)
template <typename T>
__global__ void dft_accumulate_kernel(const T* ring, i64 n_steps, i64 n_values, const double* tw, int n_freqs,
                                      double* acc) {
  const i64 total = n_values * n_freqs;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 f = t / n_values, v = t % n_values;       // consecutive threads: consecutive ring columns
    double re = acc[2 * t], im = acc[2 * t + 1];
    const double* w = tw + 2 * f;
    for (i64 s = 0; s < n_steps; ++s) {
      const double x = (double)ring[s * n_values + v];
      re = re + x * w[2 * s * n_freqs];
      im = im + x * w[2 * s * n_freqs + 1];
    }
    acc[2 * t] = re;
    acc[2 * t + 1] = im;
  }
}

// CurrentDetector.single_point_current (fdtd/detectors.py:417-461): z-current through each cell from the
// loop of H around it, averaged over the cell's z level and the one below; indices wrap like python's.
// The second `current_vector_2` is ACCUMULATED onto the first (`+=`, fdtd/detectors.py:456) as in the reference.
// `ghost`: the slab has a left neighbour, whose last H plane lies in the ghost plane just below local plane 0
// (x-sharded grids; the caller samples after that plane has arrived) -- no wrap-around along x then.
template <typename T, typename A = T>
__global__ void current_kernel(const T* Hx, const T* Hy, const i64* idx, const int* pos, int n, int Nx, int Ny,
                               int Nz, i64 plane, A dx, T* ring, T* last, i64 slot, int ghost) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const i64 lin = idx[t];
    const i64 px = lin / plane, py = (lin % plane) / Nz, pz = lin % Nz;
    const i64 pxm = (px > 0 || ghost) ? px - 1 : Nx - 1;
    const i64 pym = py > 0 ? py - 1 : Ny - 1;
    const i64 pzm = pz > 0 ? pz - 1 : Nz - 1;
    A cv1 = ((A)Hx[px * plane + pym * Nz + pz] - (A)Hx[px * plane + py * Nz + pz]) * dx;
    A cv2 = ((A)Hy[px * plane + py * Nz + pz] - (A)Hy[pxm * plane + py * Nz + pz]) * dx;
    const A c1 = cv1 + cv2;
    cv1 = ((A)Hx[px * plane + pym * Nz + pzm] - (A)Hx[px * plane + py * Nz + pzm]) * dx;
    cv2 = cv2 + ((A)Hy[px * plane + py * Nz + pzm] - (A)Hy[pxm * plane + py * Nz + pzm]) * dx;
    const A c2 = cv1 + cv2;
    const T I = (T)((c1 + c2) / A(2));
    ring[slot * n + pos[t]] = I;
    last[pos[t]] = I;
  }
}

// SoftArbitraryPointSource.update_E (fdtd/sources.py:596-626).
//   wave[w]     = input voltage of the step, in the grid dtype
//   wave_div[w] = input voltage / grid spacing evaluated in float64 on the host, then rounded: what the
//                 reference adds when no current enters (Z <= 0, or the very first step) -- there the whole
//                 expression is host float64 arithmetic
//   otherwise   vout = vin + Z * I_prev and E += vout / dx, in the grid dtype
template <typename T, typename A = T>
__global__ void source_feedback_kernel(T* F, i64 cell, const A* wave, const A* wave_div, i64 w, A impedance,
                                       const T* last_I, int use_current, A dx, T* record, i64 slot) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const A vin = wave[w];
    A vout = vin;
    if (impedance > A(0) && use_current) {
      vout = vin + impedance * (A)last_I[0];
      F[cell] = (T)((A)F[cell] + vout / dx);
    } else {
      F[cell] = (T)((A)F[cell] + wave_div[w]);
    }
    if (record) record[slot] = (T)vout;
  }
}

// ---- direct peer-to-peer halo (one process per GPU, peer pointers from CUDA IPC) --------------------------

// unfused push: copy the boundary plane of the y and z components into the neighbour's ghost planes
template <typename T>
__global__ void halo_push_kernel(const T* src_y, const T* src_z, T* dst_y, T* dst_z, i64 n) {
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (i64)gridDim.x * blockDim.x) {
    dst_y[t] = src_y[t];
    dst_z[t] = src_z[t];
  }
}

#ifndef FDTD_EMU
// publish: everything this stream wrote before (kernel boundaries order it) is visible system-wide, then the
// neighbour's flag takes the new half-step count
__global__ void halo_signal_kernel(i64* peer_flag, i64 value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(value) : "memory");
  }
}

// consume: spin until the local flag (written by the neighbour) reaches `value`.  A neighbour that stopped
// stepping must not turn into a silent wrong result: after `timeout_ns` of wall-clock time (%globaltimer) the
// kernel raises *error and TRAPS, so nothing enqueued behind it ever consumes a stale ghost plane -- the next
// synchronising call of the host reports the failure.
__global__ void halo_wait_kernel(const i64* flag, i64 value, int* error, i64 timeout_ns) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    i64 v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
      if (v >= value) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (timeout_ns > 0 && (i64)(t1 - t0) > timeout_ns) {
        *error = 1;
        __threadfence_system();
        __trap();
      }
      __nanosleep(200);
    }
  }
}
#endif

// graph replays: the per-launch bases of the waveform index and the ring slot
__global__ void set_dyn_kernel(i64* dyn, i64 wave_base, i64 slot_base) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    dyn[0] = wave_base;
    dyn[1] = slot_base;
  }
}

}  // namespace fdtd
