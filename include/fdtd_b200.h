/* fdtd_b200.h -- C ABI of the B200-native Yee-update engine.
 *
 * This is the drop-in boundary for the ONE hot path of flaport/fdtd: the
 * per-timestep update Grid.step() = update_E + update_H (reference
 * fdtd/grid.py:267-325) with everything the reference runs inside it: CPML
 * psi/phi updates and field corrections (fdtd/boundaries.py:409-487), periodic
 * copies (fdtd/boundaries.py:184-219), Object / AbsorbingObject /
 * AnisotropicObject region updates (fdtd/objects.py:118-129, 207-221,
 * 254-269), Point/Line/Plane source injection (fdtd/sources.py:93-109,
 * 278-297, 476-486) and Line/Block detector sampling
 * (fdtd/detectors.py:114-124, 241-263).
 *
 * The reference has no FFI: its seam is the Python backend singleton
 * (fdtd/backend.py:363-439) plus the duck-typed plug-in protocol Grid calls
 * every half-step (fdtd/grid.py:279-299, 305-325).  A maintainer binds this
 * library with ctypes (INTEGRATION.md shows the stub); fdtd_b200/_capi.py is
 * that binding.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch / C++ types.
 *  - every device pointer is owned by the caller (PyTorch tensors on the Python
 *    side); the library never allocates or frees field memory.
 *  - every entry point returns 0 on success or a negative FDTD_ERR_* code and
 *    never throws; fdtd_last_error() gives the thread-local message.
 *  - work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t
 *    passed as void*); nothing synchronises implicitly.
 *  - fields are SoA: one array per component, C-order [x][y][z], z contiguous,
 *    `plane` = Ny*Nz elements between consecutive x-planes.  Each component
 *    array has one ghost x-plane below local plane 0 and one above local plane
 *    Nx-1 (pointers point at local plane 0); the ghosts carry the neighbour
 *    slab's boundary plane when the grid is sharded in x and make every
 *    stencil read in-bounds otherwise.
 */
#ifndef FDTD_B200_H
#define FDTD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD_ABI_VERSION 15

#define FDTD_F32 0
#define FDTD_F64 1
#define FDTD_F32X 2   /* float storage of the STATE (fields, psi, detector rings), double arithmetic: every value is
                         widened where it is used and rounded once where it is stored; every COEFFICIENT (material
                         arrays, CPML tables, source profiles and waveform tables) is double.  The reference's
                         ".float32" backends compute in float64 (fdtd/backend.py:43-49, 281-284): this mode keeps
                         the float32 footprint and traffic but stays within 1e-5 of that reference over thousands
                         of steps, where float32 arithmetic drifts (1.6e-5 at 2000 steps) */

#define FDTD_MAX_SLABS 6
#define FDTD_MAX_POST 16
#define FDTD_FUSED_MAX 6       /* sources / detectors per field folded into the half-step kernel */

#define FDTD_OK 0
#define FDTD_ERR_ARG (-1)     /* invalid descriptor / argument */
#define FDTD_ERR_CUDA (-2)    /* a CUDA runtime call or launch failed */
#define FDTD_ERR_UNSUPPORTED (-3)

/* tile_class bits (one byte per (x-plane, y-tile, z-tile)) */
#define FDTD_CLS_VARY_E 1   /* eps^-1 differs from the background somewhere in the tile: stream inv_eps */
#define FDTD_CLS_VARY_H 2   /* mu^-1 differs from the background: stream inv_mu */
#define FDTD_CLS_ABSORB 4   /* an AbsorbingObject covers part of the tile: stream absorb */
#define FDTD_CLS_OBJECT 8   /* an Object covers part of the tile: the PML add needs inv_eps_grid */
#define FDTD_CLS_OVERLAP 32 /* two objects overlap in the tile: the second one's update is applied separately (inv_eps2,
                               absorb2), as the reference's per-object loop does (fdtd/objects.py:127-129) */
#define FDTD_CLS_ANISO 16   /* an AnisotropicObject covers part of the tile: its cells round sc*(eps^-1*curl) like the
                               reference's bmm, not (sc*eps^-1)*curl.  They are marked by a NEGATIVE zero in inv_eps_grid:
                               the x-component when it is the first object covering the cell, the y-component when it is
                               the second */
#define FDTD_CLS_ABSORB2 64 /* the second object covering some cell of the tile is an AbsorbingObject: stream absorb2 */

/* post-op kinds, executed in registration order after the fused half-step kernel */
#define FDTD_POST_PERIODIC 0  /* arg = axis: E[0]=E[-1] after E, H[-1]=H[0] after H  (fdtd/boundaries.py:184-219) */
#define FDTD_POST_PML_ADD 1   /* arg = slab index: E[loc] += sc*eps^-1*phi_E for a PML registered after a periodic boundary */

/* source kinds */
#define FDTD_SRC_POINTS 0     /* soft: F[comp][idx[n]] += profile[n] * wave[q]   (Point/LineSource) */
#define FDTD_SRC_BOX 1        /* hard: F[comp][box]     = amplitude  * wave[q]   (PlaneSource) */
#define FDTD_SRC_FEEDBACK 2   /* soft voltage source with series impedance (SoftArbitraryPointSource,
                                 fdtd/sources.py:596-626): Ez[idx[0]] += (wave[q] + Z * I_prev) / spacing, where
                                 I_prev is the last sample of a current detector, kept on the device */

/* detector kinds */
#define FDTD_DET_FIELD 0      /* E after the E half-step, H after the H half-step: [capacity][n][3] rings */
#define FDTD_DET_CURRENT 1    /* CurrentDetector (fdtd/detectors.py:417-476): after the H half-step, the z-current
                                 from the H loop around each cell (two z-levels averaged): ring_H is [capacity][n] */

/* One CPML slab (replaces the 14 full-size arrays per PML of fdtd/boundaries.py:367-407).
 * Local index l = global index along `axis` - lo.  Only two psi scalars per cell and per
 * field can ever be non-zero:
 *   psi[0] is driven by d F_w / d axis and corrects component u = axis+1 with sign -,
 *   psi[1] is driven by d F_u / d axis and corrects component w = axis+2 with sign +.
 * psi layout (psi[1] starts psi_count elements after psi[0]):
 *   axis 0: [x - x0][y][z]   for local planes x0 <= x < x1
 *   axis 1: [x][l][z]
 *   axis 2: [x][y][r], rows of length R = roundup4(lo + thickness - (lo & ~3)) whose entry r is the cell
 *           z = (lo & ~3) + r; entries outside the slab are padding (zero) -- keeps psi access 128-bit aligned */
typedef struct fdtd_slab {
  int32_t axis;
  int32_t lo;          /* first GLOBAL index of the slab along axis */
  int32_t thickness;
  int32_t fused;       /* 1: field correction applied inside the half-step kernel */
  int32_t x0, x1;      /* local x-plane range covered by this slab's psi storage */
  int64_t psi_count;   /* elements per psi scalar */
  void* psi_E;         /* device [2][psi_count] */
  void* psi_H;         /* device [2][psi_count] */
  const void* bE;      /* device [thickness]: exp(-(sigma_E/k + a) * sc)   fdtd/boundaries.py:396 */
  const void* cE;      /* device [thickness]: (bE-1)*sigma_E/(sigma_E*k + a*k^2)   :397-401 */
  const void* bH;
  const void* cH;
} fdtd_slab;

typedef struct fdtd_source {
  int32_t kind;        /* FDTD_SRC_* */
  int32_t field;       /* 0 = E (applied after the E half-step), 1 = H */
  int32_t comp;        /* component written */
  int32_t n;           /* FDTD_SRC_POINTS: number of points on this slab */
  const int64_t* idx;  /* device [n]: local linear cell index x*plane + y*Nz + z, ASCENDING */
  const void* profile; /* device [n], same order as idx */
  double amplitude;    /* FDTD_SRC_BOX */
  int32_t box[6];      /* FDTD_SRC_BOX: local x0,x1,y0,y1,z0,z1 (half-open) */
  const void* wave;    /* device [wave_len]: per-step scalar, entry q - wave_q0 */
  int64_t wave_q0;
  int64_t wave_len;
  int32_t bbox[6];     /* FDTD_SRC_POINTS: local bounding box x0,x1,y0,y1,z0,z1 (half-open) of the points */
  /* FDTD_SRC_FEEDBACK only.  wave = input voltage per step; profile = device [wave_len] input voltage /
   * spacing evaluated in float64 and rounded (added as is when no current enters the step: Z <= 0 or q = 0). */
  double impedance;    /* Z (ohm); <= 0: plain voltage source */
  double spacing;      /* grid spacing: volts -> field */
  const void* feedback;/* device [1]: last current sample of the paired detector (fdtd_detector.last) */
  void* record;        /* device [record_capacity]: output voltage per step, or NULL */
  int64_t record_capacity;
} fdtd_source;

typedef struct fdtd_detector {
  int32_t n;           /* points on this slab */
  int32_t kind;        /* FDTD_DET_* */
  const int64_t* idx;  /* device [n]: local linear cell index, ASCENDING */
  const int32_t* pos;  /* device [n]: position of each entry in the detector's sampling order (ring column) */
  void* ring_E;        /* device [capacity][n][3] */
  void* ring_H;        /* device [capacity][n][3] */
  int64_t capacity;
  int32_t bbox[6];     /* local bounding box of the points */
  void* last;          /* FDTD_DET_CURRENT: device [n], most recent sample (feeds FDTD_SRC_FEEDBACK) */
  double spacing;      /* FDTD_DET_CURRENT: grid spacing */
} fdtd_detector;

/* An object that is the THIRD (or later) one covering some cells.  The reference updates every object in registration
 * order (fdtd/grid.py:285-287), whatever the depth; the fused kernel applies the first two objects covering a cell
 * (coefficient layers 1 and 2), every further one is applied by its own small kernel over its box, in registration
 * order, right after the fused kernel -- exactly its update_E on the cells named by `mask`
 * (fdtd/objects.py:118-129, 207-221, 254-269). */
#define FDTD_OBJ_PLAIN 0
#define FDTD_OBJ_ANISO 1
#define FDTD_OBJ_ABSORB 2
typedef struct fdtd_deep_object {
  int32_t kind;           /* FDTD_OBJ_* */
  int32_t box[6];         /* local x0,x1,y0,y1,z0,z1 (half-open) */
  const void* inv[3];     /* device [x1-x0][y1-y0][z1-z0] per component: the object's eps^-1 over its box (coefficient type) */
  const void* absorb[3];  /* FDTD_OBJ_ABSORB: absorption factor over the box */
  const uint8_t* mask;    /* device [x1-x0][y1-y0][z1-z0]: 1 where two earlier objects already cover the cell */
} fdtd_deep_object;

typedef struct fdtd_desc {
  int32_t abi_version; /* FDTD_ABI_VERSION */
  int32_t dtype;       /* FDTD_F32 / FDTD_F64: storage and arithmetic type; FDTD_F32X: float state, double arithmetic
                          and coefficients -- below, "device [n]" arrays of coefficients are double in that mode */
  int32_t Nx, Ny, Nz;  /* LOCAL slab extents */
  int32_t x_offset;    /* global index of local plane 0 */
  int32_t Nx_global;
  int32_t pad0_;
  int64_t plane;       /* Ny*Nz */
  void* E[3];          /* device, local plane 0 of Ex, Ey, Ez */
  void* H[3];
  double courant;      /* sc = grid.courant_number (fdtd/grid.py:116-127) */
  double bg_inv_eps[3];   /* background eps^-1 used by tiles without FDTD_CLS_VARY_E */
  double bg_inv_mu[3];
  const void* inv_eps[3];      /* effective eps^-1 of the curl term: grid value outside objects, object value inside; or NULL */
  const void* inv_eps2[3];     /* eps^-1 of the SECOND object covering a cell (zero elsewhere), or NULL: no overlaps */
  const void* inv_eps_grid[3]; /* the grid's own eps^-1 (zero inside objects, fdtd/objects.py:92) for the PML add; NULL = inv_eps */
  const void* absorb[3];       /* AbsorbingObject absorption factor f (fdtd/objects.py:198-205), zero elsewhere; or NULL */
  const void* absorb2[3];      /* the same for the SECOND object covering a cell; or NULL */
  const void* inv_mu[3];       /* or NULL */
  const uint8_t* tile_class;   /* device [Nx][tiles_y][tiles_z], or NULL = every tile homogeneous */
  const uint8_t* plane_class;  /* HOST [Nx], optional: the OR of tile_class over each x-plane.  Runs of planes without any
                                  class bit are then launched with the material-free instantiation of the kernel
                                  (no coefficient / object code, fewer registers): objects confined to a part of
                                  the x-range cost nothing outside it */
  int32_t tile_y, tile_z;      /* tile extents in cells, as returned by fdtd_tile_shape */
  int32_t n_slabs;
  int32_t n_post;
  fdtd_slab slabs[FDTD_MAX_SLABS];        /* registration order */
  int32_t post_kind[FDTD_MAX_POST];
  int32_t post_arg[FDTD_MAX_POST];
  int32_t n_sources;
  int32_t n_detectors;
  const fdtd_source* sources;      /* HOST array [n_sources], registration order; any length (the reference keeps plain
                                      Python lists, fdtd/grid.py:155-163); caller-owned, read during the call only */
  const fdtd_detector* detectors;  /* HOST array [n_detectors], registration order */
  int32_t x_chunk;     /* planes marched per thread block; 0 = library default */
  int32_t use_graphs;  /* 1: fdtd_run may replay CUDA graphs of step chunks (launch-bound small grids) */
  int64_t* dyn;        /* device int64[2] scratch owned by the caller, needed when use_graphs = 1 */
  int32_t fuse_eh;     /* != 0: fdtd_run may run pairs of temporally fused E+H steps -- one kernel per step that moves 12
                          instead of 18 words per cell -- on homogeneous unsharded grids without periodic boundaries,
                          with point sources on E only; needs E2 / H2 / psi_E2.  1: wherever that is legal,
                          2: only where it is also faster than the two half-steps (y-z planes of 1.07 MiB per component and
                          more whose z extent fills the kernel's 31-vector tiles to 85 %, 64 x-planes and more; CPML slabs of
                          up to 32 cells in either mode) */
  int32_t pad2_;
  void* E2[3];         /* second field buffers of the ping-pong pair, same layout as E / H (ghost planes included), */
  void* H2[3];         /* or NULL; after fdtd_run the results are always in E / H */
  int32_t fuse_post;   /* sources/detectors folded into the half-step kernel: 1 always (when legal), 0 never,
                          -1 automatic (local slabs up to 2^25 cells, where the launches between the half-steps still show) */
  int32_t pad3_;
  void* psi_E2[FDTD_MAX_SLABS];  /* fuse_eh: a second psi_E buffer [2][psi_count] for every slab.  The fused kernel updates the
                          whole grid, CPML cells and faces included, in its single pass; cells whose E_new is recomputed
                          by a neighbouring thread need the OLD psi_E, so psi_E alternates between the two buffers like
                          the fields (results always end in psi_E) */
  int32_t n_deep;      /* objects beyond the second one on a cell */
  int32_t h_wrap_ghost;/* x-sharded grids: 1 = the caller keeps the LOW ghost plane of Hy on the first slab equal to the last
                          slab's last plane, so that a CurrentDetector cell on global plane x = 0 finds H[x-1] = H[-1]
                          there (fdtd/detectors.py:432-447 wraps like python indexing) */
  const fdtd_deep_object* deep;  /* HOST array [n_deep], registration order */
  int32_t x_wrap;      /* x-sharded grid with a periodic x boundary (fdtd/boundaries.py:184-195): 0 = none, else
                          1 + the number of post ops registered before it.  The copy E[0] = E[-1] / H[-1] = H[0] then
                          crosses the first and last slab: the caller runs fdtd_post_part(.., 0, ..), moves the plane
                          between the two ranks itself, then fdtd_post_part(.., 1, ..) */
  int32_t pad4_;
} fdtd_desc;

/* --- queries ------------------------------------------------------------------------- */
int32_t fdtd_abi_version(void);
/* sizeof(fdtd_desc) as the library was compiled: a binding checks its own struct layout against it */
int64_t fdtd_sizeof_desc(void);
int64_t fdtd_sizeof_halo(void);
const char* fdtd_last_error(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t fdtd_launch_count(void);
/* tile extents the half-step kernels use for a (dtype, Ny, Nz) grid; tile_class must be laid out with them */
int fdtd_tile_shape(int32_t dtype, int32_t Ny, int32_t Nz, int32_t* tile_y, int32_t* tile_z);
/* validate a descriptor without launching anything */
int fdtd_validate(const fdtd_desc* d);

/* --- the hot path -------------------------------------------------------------------- */
/* Fused E half-step on local planes [x_begin, x_end): psi_E update of every slab, curl_H,
 * E += sc*eps^-1*curl with object / absorbing coefficients, fused PML field corrections and --
 * when fdtd_post_is_fused(d) -- source injection (step index q) and detector sampling (ring slot
 * `slot`) in the same pass.
 * Replaces PML.update_phi_E + curl_H + grid.py:283 + Object.update_E + PML.update_E
 * (+ Source.update_E + Detector.detect_E). */
int fdtd_e_halfstep(const fdtd_desc* d, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot, void* stream);
/* Same for H: PML.update_phi_H + curl_E + grid.py:309 + PML.update_H (+ sources, detectors). */
int fdtd_h_halfstep(const fdtd_desc* d, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot, void* stream);
/* 1 when sources and detectors are folded into the half-step kernels: d->fuse_post allows it, nothing
 * has to run between the field update and them (no periodic copy, no late PML correction) and there
 * are at most FDTD_FUSED_MAX of each per field.  fdtd_post_E/H are then no-ops. */
int fdtd_post_is_fused(const fdtd_desc* d);
/* What follows the fused kernel inside Grid.update_E / update_H when it could not be folded, in
 * the reference's order: periodic copies and late PML corrections (registration order), sources
 * (registration order, step index q), detector sampling into ring slot `slot`. */
int fdtd_post_E(const fdtd_desc* d, int64_t q, int64_t slot, void* stream);
int fdtd_post_H(const fdtd_desc* d, int64_t q, int64_t slot, void* stream);
/* The same in two parts around the plane transfer of an x-sharded periodic x boundary (d->x_wrap != 0):
 * part 0 = the post ops registered before that boundary, part 1 = the remaining post ops, sources, detectors.
 * field: 0 = E, 1 = H. */
int fdtd_post_part(const fdtd_desc* d, int32_t field, int32_t part, int64_t q, int64_t slot, void* stream);
/* The general form: any subset of the phases that follow the half-step kernel, in their fixed order.  A caller that
 * has to do something in between (move the wrap plane of a periodic x boundary; exchange the ghost planes BEFORE the
 * detectors sample, for a CurrentDetector whose H loop reaches into the neighbour slab) calls it more than once. */
#define FDTD_PHASE_OBJECTS 16u   /* objects beyond the second one on a cell (fdtd_deep_object), registration order */
#define FDTD_PHASE_BEFORE 1u     /* the boundary post ops registered before the x-wrap boundary (all of them when
                                    d->x_wrap == 0) */
#define FDTD_PHASE_AFTER 2u      /* the boundary post ops registered after the x-wrap boundary */
#define FDTD_PHASE_SOURCES 4u
#define FDTD_PHASE_DETECTORS 8u
#define FDTD_PHASE_ALL 31u
int fdtd_post_phases(const fdtd_desc* d, int32_t field, uint32_t phases, int64_t q, int64_t slot, void* stream);
/* Grid.update_E / Grid.update_H (fdtd/grid.py:275-325) on the whole local slab */
int fdtd_update_E(const fdtd_desc* d, int64_t q, int64_t slot, void* stream);
int fdtd_update_H(const fdtd_desc* d, int64_t q, int64_t slot, void* stream);
/* Grid.run (fdtd/grid.py:250-265): nsteps full steps starting at step index q0; detector
 * samples of step q0+s go to ring slot slot0+s */
int fdtd_run(const fdtd_desc* d, int64_t q0, int64_t nsteps, int64_t slot0, void* stream);

/* 1 when fdtd_run will execute pairs of temporally fused E+H steps for this descriptor (fuse_eh set, second
 * buffers given, homogeneous unsharded grid, no periodic boundary, point sources on E only) */
int fdtd_fuse_eh_active(const fdtd_desc* d);

/* --- direct peer-to-peer halo exchange (x-sharded grids, one process per GPU) -------------------------
 * No reference counterpart (the reference is single-device).  Each rank exports its field storage and a
 * two-word flag array with CUDA IPC; neighbours import them and get peer pointers that the kernels below
 * store through over NVLink.  Flags carry monotonically increasing half-step counts, so no host-side
 * ordering between the processes is needed. */
/* IPC handle (64 bytes) of the allocation holding dev_ptr and dev_ptr's offset inside it */
int fdtd_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset);
/* map a neighbour's allocation (lazy peer access) and return the pointer at `offset` */
int fdtd_ipc_import(const void* handle64, int64_t offset, void** dev_ptr);
/* fdtd_e_halfstep / fdtd_h_halfstep (field 0 / 1) on a plane range that contains the slab's boundary plane
 * (local plane 0 for E, Nx-1 for H): the threads that compute that plane also store its y and z components
 * into the neighbour's ghost planes -- compute and halo transfer in one kernel */
int fdtd_halfstep_push(const fdtd_desc* d, int32_t field, int32_t x_begin, int32_t x_end, int64_t q, int64_t slot,
                       void* peer_ghost_y, void* peer_ghost_z, void* stream);
/* the same transfer as a separate copy kernel, for steps where something modifies the boundary plane after
 * the half-step kernel (periodic copies, sources on that plane) */
int fdtd_halo_push(const fdtd_desc* d, int32_t field, void* peer_ghost_y, void* peer_ghost_z, void* stream);
/* publish `value` in the neighbour's flag after everything enqueued before on `stream` (release, system scope) */
int fdtd_halo_signal(int64_t* peer_flag, int64_t value, void* stream);
/* make `stream` wait until the local flag reaches `value` (acquire, system scope).  After `timeout_ns` of wall-clock
 * time (0 = wait for ever) the waiting kernel sets *error (device int) and TRAPS: nothing enqueued behind it runs on
 * a stale ghost plane, and the caller's next synchronising CUDA call fails */
int fdtd_halo_wait(const int64_t* flag, int64_t value, int32_t* error, int64_t timeout_ns, void* stream);

/* One rank's view of the peer-to-peer halo exchange: peer pointers into the two neighbour slabs (from
 * fdtd_ipc_import), this rank's flag words, and the running push counts.  Caller-owned host struct; `count` is
 * updated by the calls below. */
typedef struct fdtd_halo {
  int32_t has_left, has_right;   /* this slab has a left / right neighbour */
  void* left_ghost_y;            /* LEFT neighbour's HIGH ghost plane of Ey (peer pointer) */
  void* left_ghost_z;            /*   ... of Ez: where this slab's E plane 0 goes after every E half-step */
  void* right_ghost_y;           /* RIGHT neighbour's LOW ghost plane of Hy */
  void* right_ghost_z;           /*   ... of Hz: where this slab's last H plane goes after every H half-step */
  void* left_ghost_y2;           /* the same four ghost planes in the neighbours' SECOND field buffers (fdtd_desc E2 / H2): */
  void* left_ghost_z2;           /*   a temporally fused step writes the other buffer of the ping-pong pair, and its    */
  void* right_ghost_y2;          /*   boundary planes go into the neighbour's buffer of the same parity.  NULL: no      */
  void* right_ghost_z2;          /*   fused steps on this slab                                                          */
  int64_t* left_flag;            /* left neighbour's flag word [0]: number of E pushes it has received */
  int64_t* right_flag;           /* right neighbour's flag word [1]: number of H pushes it has received */
  const int64_t* flags;          /* this rank's device int64[2], written by the neighbours */
  int32_t* error;                /* device int32, raised by a wait that timed out */
  int64_t count[2];              /* E / H pushes done so far == pushes expected from the neighbours (in / out) */
  int32_t push_fused[2];         /* E / H: nothing modifies the boundary plane after the half-step kernel (no periodic
                                    copy, late PML correction or unfused source on it): the kernel that computes the
                                    plane stores it into the neighbour's ghost as well; otherwise a copy kernel
                                    pushes it after the post ops */
  void* side_stream;             /* cudaStream_t the boundary plane runs on, concurrently with the bulk */
  int64_t timeout_ns;            /* fdtd_halo_wait limit; 0 = none */
} fdtd_halo;

/* One half-step (field 0 = E, 1 = H) of an x-sharded slab, everything a rank does for it: the bulk planes on
 * `stream`; on h->side_stream the wait for the neighbour's ghost plane, the boundary plane (pushed into the
 * neighbour's ghost by the same kernel when h->push_fused allows) and the flag that publishes it; then the post
 * ops, sources and detectors (fdtd_post_E/H) on `stream`; the streams are joined with events.  Not for slabs with
 * d->x_wrap (the caller moves the wrap plane between fdtd_post_part calls itself). */
int fdtd_sharded_halfstep(const fdtd_desc* d, fdtd_halo* h, int32_t field, int64_t q, int64_t slot, void* stream);
/* Grid.run on an x-sharded slab: nsteps x { E half-step, H half-step } as above, no host work in between --
 * the ranks only meet through the flag words */
int fdtd_run_sharded(const fdtd_desc* d, fdtd_halo* h, int64_t q0, int64_t nsteps, int64_t slot0, void* stream);
/* 1 when fdtd_run_sharded will execute pairs of temporally fused E+H steps on this slab (d->fuse_eh, second buffers
 * and their peer ghosts given, homogeneous grid, no periodic boundary, point sources on E only): per step one fused
 * kernel that also stores E_new[plane 0] into the left neighbour's ghost, then -- once the right neighbour's E_new has
 * arrived -- the H update of the last plane, stored into the right neighbour's ghost as well */
int fdtd_fuse_eh_sharded_active(const fdtd_desc* d, const fdtd_halo* h);
/* push both boundary planes and wait for the neighbours' (collective: every rank calls it; after the user wrote
 * E / H, and once after set-up) */
int fdtd_halo_refresh(const fdtd_desc* d, fdtd_halo* h, void* stream);

/* ---- running DFT of detector records, on the device (spectra of long runs without
 * moving the time traces to the host; the reference transforms the host lists, fdtd/fourier.py:172-213) -----
 * ring     device [n_steps][n_values] samples of the library dtype (a filled part of a detector ring)
 * twiddle  device double [n_steps][n_freqs][2]: (cos, sin) of -2 pi f n dt for the record index n of each row,
 *          tabulated by the host like the source waveforms
 * acc      device double [n_freqs][n_values][2], updated: acc[f][v] += sum_s ring[s][v] * twiddle[s][f]
 *          (rows added in order, in float64, so the result does not depend on how a record is split into calls) */
int fdtd_dft_accumulate(int32_t dtype, const void* ring, int64_t n_steps, int64_t n_values, const double* twiddle,
                        int32_t n_freqs, double* acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FDTD_B200_H */
