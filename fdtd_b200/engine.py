"""Host side of the fused Yee update: bakes a Grid's registrations into the C-ABI descriptor
(include/fdtd_b200.h) and drives the CUDA kernels.

What is baked (once, and again only when something registered or a material array changed):
  * CPML slabs (registration order) with their psi storage and 1-D b/c tables; slabs registered
    before the first periodic boundary are corrected inside the fused kernel, later ones by a
    post-op so that the reference's registration-order semantics hold (fdtd/grid.py:290-291);
  * materials: the grid's eps^-1 / mu^-1 arrays only if they exist (they are created lazily --
    a homogeneous grid streams no coefficient array at all), the effective eps^-1 of the curl
    term (grid + objects), the absorption factor, and the per-(plane, tile) class byte;
  * sources (point lists / hard boxes) and their host-tabulated waveforms;
  * detector point lists and device ring buffers, flushed to the host in batches.
"""
import ctypes as C

import torch

from . import _capi
from .backend import backend as bd
from .sharding import HaloExchange, P2PHalo, WrapExchange

RING_BYTES = 64 << 20          # detector ring budget per grid
WAVE_TABLE_MIN = 1024          # look-ahead of the host waveform tables (steps)
WAVE_TABLE_MAX = 1 << 16
GRAPH_MAX_CELLS = 1 << 25      # grids up to this many cells replay CUDA graphs of 32-step chunks in run()
# automatic mode of the single-pass E+H kernel (grid._fuse_eh = 2): where it beats the two half-steps on the B200
# (profiles/r2_s14/fused_sizes.log: float32 512^3 -5 %, 576^3 +12 %, 640^3 +7 %, 768^3 +13 %, 1024^3 +25 %; float64
# 384^3 +6 %, 512^3 +19 %; slabs of 1024^2 planes: 64 planes +2 %, 128 +10 %, 256 +17 %) -- the same test as
# fuse_eh_eligible() in csrc/fdtd_b200.cu
FUSE_EH_MIN_PLANE_BYTES = 1100 << 10  # y-z plane of one component
FUSE_EH_MIN_PLANES = 64               # x-planes (unsharded)
FUSE_EH_MIN_Z_FILL = 0.85             # Nz / (z tiles of 31 vectors x their length)
FUSE_EH_MAX_PML = 32                  # cells per CPML slab (the fused kernel keeps the tables in shared memory)
FUSE_EH_MIN_SLAB = 96                 # x-planes per rank (x-sharded; 128 planes per rank: 1.41 against 1.64 ms per step, profiles/r2_n2c/)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(None)


class Engine:
    def __init__(self, grid):
        self.grid = grid
        self.lib = bd.lib
        self.desc = _capi.Desc()
        self._keep = []            # tensors referenced by raw pointers in the descriptor
        self._wave = None          # (q0, len)
        self._halo = None
        self._p2p = None
        self._wrap = None
        self._late = False
        self._pending = {"E": None, "H": None}
        self.bake()

    # ------------------------------------------------------------------------------------ bake
    def bake(self):
        g, d = self.grid, self.desc
        part = g._part
        dt = g._dtype
        C.memset(C.byref(d), 0, C.sizeof(d))
        self._keep = []
        d.abi_version = _capi.ABI_VERSION
        d.dtype = (_capi.F32 if dt is torch.float32 else
                   _capi.F32X if g._sdtype is torch.float32 else _capi.F64)
        d.Nx, d.Ny, d.Nz = part.nx, g.Ny, g.Nz
        d.x_offset, d.Nx_global = part.x0, g.Nx
        d.plane = g.Ny * g.Nz
        d.courant = g.courant_number
        d.x_chunk = g._x_chunk
        for c in range(3):
            d.E[c] = g._E[c, 1].data_ptr()
            d.H[c] = g._H[c, 1].data_ptr()
            d.bg_inv_eps[c] = g._bg_inv_eps[c]
            d.bg_inv_mu[c] = g._bg_inv_mu[c]

        # --- plug-ins -------------------------------------------------------------------------
        # Built-in plug-ins are folded into the kernels.  Anything else registered on the grid -- a class of the user's
        # that follows the reference's duck-typed protocol (fdtd/grid.py:279-299, 305-325), or a subclass of a built-in
        # that overrides a protocol method -- is called from Python on every half-step, at the reference's place in
        # the order (`_hooked_halfstep`): a documented slow path, never a silent skip.
        from .boundaries import PML, PeriodicBoundary
        from .detectors import _Detector
        from .objects import Object

        def custom(thing, *methods):
            return [m for m in methods
                    if callable(getattr(thing, m, None))
                    and not getattr(getattr(type(thing), m, None), "_fdtd_b200_builtin", False)]

        self._hooks = {
            "boundaries": [(b, custom(b, "update_phi_E", "update_phi_H", "update_E", "update_H")) for b in g.boundaries],
            "objects": [(o, custom(o, "update_E", "update_H")) for o in g.objects],
            "sources": [(x, custom(x, "update_E", "update_H")) for x in g.sources],
            "detectors": [(x, custom(x, "detect_E", "detect_H")) for x in g.detectors],
        }
        self._hooks = {k: [(t, m) for t, m in v if m] for k, v in self._hooks.items()}
        self._hooked = any(self._hooks.values())
        if self._hooked and part.sharded:
            raise NotImplementedError("user-defined plug-ins (objects / sources / detectors / boundaries with their own "
                                      "update or detect methods) on an x-sharded grid: they see one slab only")
        objects = [o for o in g.objects if isinstance(o, Object)]
        sources = [x for x in g.sources if hasattr(x, "_entries")]
        self._dets = [x for x in g.detectors if isinstance(x, _Detector)]

        # --- boundaries -----------------------------------------------------------------------
        slabs, post, seen_periodic, x_wrap = [], [], False, 0
        # a user object's update_E runs between the field update and the PML corrections: keep those out of the kernel
        unfuse = bool(self._hooks["objects"])
        for b in g.boundaries:
            if isinstance(b, PML):
                if len(slabs) == _capi.MAX_SLABS:
                    raise ValueError("at most six PMLs (one per face)")
                idx = len(slabs)
                slabs.append(b)
                s = d.slabs[idx]
                s.axis, s.lo, s.thickness = b.axis, b.lo, b.thickness
                s.fused = 0 if (seen_periodic or unfuse) else 1
                s.x0, s.x1 = b._x0, b._x1
                s.psi_count = b._psi_E.shape[1]
                s.psi_E, s.psi_H = _ptr(b._psi_E), _ptr(b._psi_H)
                s.bE, s.cE, s.bH, s.cH = (_ptr(b._tab[k]) for k in ("bE", "cE", "bH", "cH"))
                if seen_periodic or unfuse:
                    post.append((_capi.POST_PML_ADD, idx))
            elif isinstance(b, PeriodicBoundary):
                if (g.Nx, g.Ny, g.Nz)[b.axis] < 2:
                    continue                      # E[0] = E[-1] on a one-cell axis is the identity
                seen_periodic = True
                if b.axis == 0 and part.sharded:
                    # the copy crosses the first and the last slab: the host moves the plane between the post ops
                    # registered before and after this boundary (include/fdtd_b200.h, x_wrap)
                    x_wrap = len(post) + 1
                    continue
                post.append((_capi.POST_PERIODIC, b.axis))
            elif not any(t is b for t, _ in self._hooks["boundaries"]):
                raise TypeError(f"unsupported boundary {b!r}: neither a built-in one nor a plug-in with update methods")
        if len(post) > _capi.MAX_POST:
            raise ValueError("too many boundary post-ops")
        d.n_slabs, d.n_post, d.x_wrap = len(slabs), len(post), x_wrap
        for n, (kind, arg) in enumerate(post):
            d.post_kind[n], d.post_arg[n] = kind, arg

        # --- materials ------------------------------------------------------------------------
        ie_grid, imu = g._inv_eps, g._inv_mu
        ie_eff, absorb, ie2, absorb2 = ie_grid, None, None, None
        deep = []                    # (object, mask): objects that are the third or later one on some cells
        if objects:
            # The reference updates every object in registration order (fdtd/grid.py:285-287), so a cell covered
            # by several objects gets several updates.  The first object covering a cell becomes coefficient layer 1
            # (ie_eff / absorb), the second one layer 2 (ie2 / absorb2), both applied inside the fused kernel; every
            # further one is applied on those cells by its own kernel right after it (fdtd_deep_object).  Anisotropic
            # layers are marked by a negative zero in the grid's eps^-1 (x-component: layer 1, y: layer 2).
            from .objects import AbsorbingObject, AnisotropicObject
            ie_eff = ie_grid.clone()
            boxes = [(o.x, o.y, o.z) for o in objects]
            overlaps = any(all(max(p.start, q.start) < min(p.stop, q.stop) for p, q in zip(boxes[a], boxes[b]))
                           for a in range(len(boxes)) for b in range(a + 1, len(boxes)))
            cover = torch.zeros(ie_grid.shape[1:], dtype=torch.int8, device=ie_grid.device) if overlaps else None

            def mark(comp, loc, cells, aniso):
                region = ie_grid[comp][loc]
                sel = (region == 0) if cells is None else (cells & (region == 0))
                ie_grid[comp][loc] = torch.where(sel, torch.full_like(region, -0.0 if aniso else 0.0), region)

            for o in objects:
                if o._nx_local == 0:
                    continue
                loc = (slice(None),) + o._loc
                aniso = isinstance(o, AnisotropicObject)
                if cover is None:
                    ie_eff[loc] += o._inv_eps_soa
                    mark(0, o._loc, None, aniso)
                    if o._absorb_soa is not None:
                        if absorb is None:
                            absorb = torch.zeros_like(ie_grid)
                        absorb[loc] = o._absorb_soa
                    continue
                depth = cover[o._loc]
                first, second, later = depth == 0, depth == 1, depth >= 2
                zero = torch.zeros_like(o._inv_eps_soa)
                ie_eff[loc] += torch.where(first.unsqueeze(0), o._inv_eps_soa, zero)
                mark(0, o._loc, first, aniso)
                if bool(second.any()):
                    if ie2 is None:
                        ie2 = torch.zeros_like(ie_grid)
                    ie2[loc] += torch.where(second.unsqueeze(0), o._inv_eps_soa, zero)
                    mark(1, o._loc, second, aniso)
                if o._absorb_soa is not None:
                    if bool(first.any()):
                        if absorb is None:
                            absorb = torch.zeros_like(ie_grid)
                        absorb[loc] = torch.where(first.unsqueeze(0), o._absorb_soa, absorb[loc])
                    if bool(second.any()):
                        if absorb2 is None:
                            absorb2 = torch.zeros_like(ie_grid)
                        absorb2[loc] = torch.where(second.unsqueeze(0), o._absorb_soa, absorb2[loc])
                if bool(later.any()):
                    deep.append((o, later.to(torch.uint8).contiguous()))
                cover[o._loc] = torch.clamp(depth + 1, max=3)
            self._deep_table = (_capi.DeepObject * max(1, len(deep)))()
            d.deep = C.cast(self._deep_table, C.POINTER(_capi.DeepObject))
            for n, (o, mask) in enumerate(deep):
                e = d.deep[n]
                e.kind = (_capi.OBJ_ANISO if isinstance(o, AnisotropicObject) else
                          _capi.OBJ_ABSORB if isinstance(o, AbsorbingObject) else _capi.OBJ_PLAIN)
                box = (o._loc[0].start, o._loc[0].stop, o.y.start, o.y.stop, o.z.start, o.z.stop)
                for k in range(6):
                    e.box[k] = box[k]
                for c in range(3):
                    e.inv[c] = o._inv_eps_soa[c].data_ptr()
                    e.absorb[c] = o._absorb_soa[c].data_ptr() if o._absorb_soa is not None else None
                e.mask = mask.data_ptr()
                self._keep.append(mask)
        d.n_deep = len(deep)
        self._mat_versions = (None if ie_grid is None else ie_grid._version,
                              None if imu is None else imu._version)
        ty, tz = C.c_int32(), C.c_int32()
        _capi.check(self.lib, self.lib.fdtd_tile_shape(d.dtype, g.Ny, g.Nz, C.byref(ty), C.byref(tz)))
        d.tile_y, d.tile_z = ty.value, tz.value
        cls = None
        if ie_eff is not None or imu is not None:
            cls = self._classify(ie_eff, ie_grid if objects else None, absorb, imu, ty.value, tz.value, ie2, absorb2)
        for c in range(3):
            d.inv_eps[c] = ie_eff[c].data_ptr() if ie_eff is not None else None
            d.inv_eps2[c] = ie2[c].data_ptr() if ie2 is not None else None
            d.inv_eps_grid[c] = ie_grid[c].data_ptr() if (objects and ie_grid is not None) else None
            d.absorb[c] = absorb[c].data_ptr() if absorb is not None else None
            d.absorb2[c] = absorb2[c].data_ptr() if absorb2 is not None else None
            d.inv_mu[c] = imu[c].data_ptr() if imu is not None else None
        d.tile_class = _ptr(cls)
        self._plane_class = None
        if cls is not None and cls.numel() > 0:
            # which x-planes carry any material class at all (host): the others run the material-free kernel
            self._plane_class = cls.view(cls.shape[0], -1).amax(dim=1).to("cpu").numpy().copy()
            d.plane_class = self._plane_class.ctypes.data
        self._keep += [ie_eff, absorb, cls, ie2, absorb2]
        self.tile_class = cls

        # --- sources ----------------------------------------------------------------------------
        self._src_entries = []       # (desc index, source object)
        self._feedback_sources = []  # those that record their voltages in a device ring
        entries = [(s, entry) for s in sources for entry in s._entries()]
        # host tables of any length (the reference keeps plain lists, fdtd/grid.py:155-163); the descriptor points at them
        self._src_table = (_capi.Source * max(1, len(entries)))()
        d.sources = C.cast(self._src_table, C.POINTER(_capi.Source))
        n = 0
        for s, entry in entries:
            e = d.sources[n]
            e.kind, e.field, e.comp = entry["kind"], entry["field"], entry["comp"]
            if entry["kind"] == _capi.SRC_FEEDBACK:
                e.n = entry["n"]
                for k in range(6):
                    e.box[k] = entry["box"][k]
                e.impedance, e.spacing = entry["impedance"], g.grid_spacing
                e.feedback = _ptr(entry["feedback"])
                self._feedback_sources.append((n, s))
            elif entry["kind"] == _capi.SRC_POINTS:
                e.n = int(entry["idx"].numel())
                e.idx, e.profile = _ptr(entry["idx"]), _ptr(entry["profile"])
                for k in range(6):
                    e.bbox[k] = entry["bbox"][k]
                self._keep += [entry["idx"], entry["profile"]]
            else:
                e.amplitude = entry["amplitude"]
                for k in range(6):
                    e.box[k] = entry["box"][k]
            self._src_entries.append((n, s))
            n += 1
        d.n_sources = n
        self._wave = None
        self._src_sig = self._source_signature()

        # --- detectors --------------------------------------------------------------------------
        self._det_table = (_capi.Detector * max(1, len(self._dets)))()
        d.detectors = C.cast(self._det_table, C.POINTER(_capi.Detector))
        w = 4 if g._sdtype is torch.float32 else 8
        # the capacity must be the SAME on every rank of an x-sharded grid (a ring flush is collective): size it
        # from the largest per-rank share of every detector, which every rank can compute from the partition
        per_step = sum(2 * det._width * w * max(1, det._n_ring) for det in self._dets)
        self.ring_capacity = int(min(8192, max(16, RING_BYTES // max(1, per_step)))) if self._dets else 1 << 62
        for n, det in enumerate(self._dets):
            det._ensure_ring(self.ring_capacity)
            e = d.detectors[n]
            e.n = det._n_local
            e.kind = det._kind
            if det._kind == _capi.DET_CURRENT:
                e.last, e.spacing = _ptr(det._last), g.grid_spacing
            e.idx, e.pos = _ptr(det._idx), _ptr(det._pos)
            e.ring_E, e.ring_H = _ptr(det._ring_E), _ptr(det._ring_H)
            e.capacity = self.ring_capacity
            for k in range(6):
                e.bbox[k] = det._bbox[k]
        d.n_detectors = len(self._dets)
        for idx, src in self._feedback_sources:
            src._ensure_ring(self.ring_capacity if self._dets else 4096)
            d.sources[idx].record = _ptr(src._ring_V)
            d.sources[idx].record_capacity = src._capacity
        if self._feedback_sources and not self._dets:
            self.ring_capacity = 4096

        # temporally fused E+H steps (12 instead of 18 words per cell and step) need a second pair of field buffers and
        # a second psi_E per slab: homogeneous, unsharded grids only.  grid._fuse_eh / FDTD_B200_FUSE_EH: 0 never,
        # 1 wherever legal, 2 (default) where it is also faster -- large grids -- and the buffers fit in free memory
        want = g._fuse_eh
        tile_z = 31 * (16 // g._E.element_size())
        big = (g.Ny * g.Nz * g._E.element_size() >= FUSE_EH_MIN_PLANE_BYTES
               and g.Nz >= FUSE_EH_MIN_Z_FILL * (-(-g.Nz // tile_z) * tile_z)
               and (g.Nx // part.world >= FUSE_EH_MIN_SLAB if part.sharded else g.Nx >= FUSE_EH_MIN_PLANES))
        ok = bool(want and not self._hooked and ie_eff is None and imu is None and not post and not x_wrap
                  and g._sdtype is g._dtype and (want == 1 or big)
                  and all(b.thickness <= FUSE_EH_MAX_PML for b in slabs)
                  and all(d.sources[k].kind == _capi.SRC_POINTS and d.sources[k].field == 0 for k in range(d.n_sources))
                  and all(det._kind == _capi.DET_FIELD for det in self._dets) and d.n_sources <= _capi.FUSED_MAX)
        if part.sharded:
            # x-slabs: peer-to-peer halo only (the boundary planes go into the neighbours' second buffers), and every
            # rank must decide alike -- free memory included.  Thin slabs keep the two half-steps: a fused step ends in
            # a serial tail (flag, last H plane, flag) that the bulk of a two-pass step hides (FUSE_EH_MIN_SLAB)
            import os
            ok = ok and g._E.is_cuda and os.environ.get("FDTD_B200_HALO", "p2p") == "p2p" and self._p2p is not False
        if ok and g._E2 is None:
            ok = want == 1 or self._room_for(2 * g._E.numel() * g._E.element_size())
        if part.sharded and g._E.is_cuda:
            import torch.distributed as dist
            flag = torch.tensor([int(ok)], device=g._E.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if ok:
            if g._E2 is None:
                g._E2, g._H2 = torch.zeros_like(g._E), torch.zeros_like(g._H)
            for c in range(3):
                d.E2[c] = g._E2[c, 1].data_ptr()
                d.H2[c] = g._H2[c, 1].data_ptr()
            d.fuse_eh = 1 if (want == 1 or part.sharded) else 2      # (sharded: decided here, collectively)
            for idx, b in enumerate(slabs):          # psi_E ping-pong (include/fdtd_b200.h, psi_E2)
                if getattr(b, "_psi_E2", None) is None:
                    b._psi_E2 = torch.zeros_like(b._psi_E)
                d.psi_E2[idx] = _ptr(b._psi_E2)

        # CUDA-graph replay of step chunks pays off where a step is launch-bound (small grids)
        self._dyn = torch.zeros(2, dtype=torch.int64, device=g._E.device)
        d.dyn = _ptr(self._dyn)
        d.fuse_post = 0 if self._hooked else (-1 if g._fuse_post is None else int(bool(g._fuse_post)))
        d.use_graphs = 1 if (g._E.is_cuda and not part.sharded and not self._hooked and g._use_graphs is not False
                             and (g._use_graphs or part.nx * g.Ny * g.Nz <= GRAPH_MAX_CELLS)) else 0

        self._ensure_wave(g.time_steps_passed, 1)
        _capi.check(self.lib, self.lib.fdtd_validate(C.byref(d)))

        if part.sharded:
            self._setup_halo()
            self._wrap = WrapExchange(part, g._E, g._H) if x_wrap else None
        g._baked_counts = g._registration_count

    def _room_for(self, nbytes):
        """is there free device memory for `nbytes` more, with a margin? (the fused E+H path is optional)"""
        dev = self.grid._E.device
        if dev.type != "cuda":
            return True
        free, _ = torch.cuda.mem_get_info(dev)
        return free > nbytes + (4 << 30)

    def _setup_halo(self):
        """x-sharded grids: direct peer-to-peer ghost-plane stores (default on CUDA), or NCCL / gloo send-recv
        (FDTD_B200_HALO=nccl, CPU tests, or when CUDA IPC is not available between the ranks)."""
        import os
        g, d = self.grid, self.desc
        want = os.environ.get("FDTD_B200_HALO", "p2p")
        # detectors that read the ghost plane of the half-step being sampled: exchange first, then sample -- done on
        # the send / recv path only
        self._late = any(getattr(det, "_needs_ghost", False) for det in self._dets)
        wrap_H = any(getattr(det, "_needs_wrap", False) for det in self._dets)
        d.h_wrap_ghost = int(wrap_H and g._part.rank == 0)
        if self._late:
            want, self._p2p = "nccl", False
        if self._p2p is None and g._E.is_cuda and want == "p2p":
            try:
                self._p2p = P2PHalo(g._part, g._E, g._H, self.lib, g._E2 if d.fuse_eh else None,
                                    g._H2 if d.fuse_eh else None)
            except Exception as exc:                          # IPC refused (e.g. no peer access): NCCL path
                import warnings
                warnings.warn(f"fdtd_b200: peer-to-peer halo unavailable ({exc}); using NCCL send/recv")
                self._p2p = False
        self._halo = self._p2p if self._p2p else HaloExchange(g._part, g._E, g._H, wrap_H=wrap_H)
        # may the boundary plane be pushed by the half-step kernel itself?  Only if nothing modifies that
        # plane afterwards: no periodic copy / late PML correction, no unfused source on the plane
        n = d.Nx
        fused = bool(self.lib.fdtd_post_is_fused(C.byref(d)))
        self._push_fused = {}
        for field, plane in (("E", 0), ("H", n - 1)):
            ok = fused or (d.n_post == 0 and not (field == "E" and d.n_deep))
            if ok and not fused:
                for k in range(d.n_sources):
                    s = d.sources[k]
                    if s.field != (0 if field == "E" else 1):
                        continue
                    box = s.bbox if s.kind == _capi.SRC_POINTS else s.box
                    empty = (s.kind == _capi.SRC_POINTS and s.n == 0) or box[0] >= box[1]
                    if not empty and box[0] <= plane < box[1]:
                        ok = False
            self._push_fused[field] = ok
        if self._p2p:
            self._p2p.h.push_fused[0] = int(self._push_fused["E"])
            self._p2p.h.push_fused[1] = int(self._push_fused["H"])
            self._p2p_refresh()
        else:
            self._halo.refresh()

    def _post(self, field, q, slot, st, phases=_capi.PHASE_ALL):
        """what follows the half-step kernel on an x-sharded slab: deep objects and post ops, sources, detectors
        (`phases`: a subset, include/fdtd_b200.h FDTD_PHASE_*) -- around the plane transfer of a periodic x boundary
        if there is one."""
        lib, d = self.lib, self.desc
        fidx = 0 if field == "E" else 1
        if self._wrap is None:
            _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, phases, q, slot, st))
            return
        first = _capi.PHASE_OBJECTS | _capi.PHASE_BEFORE
        if phases & first:
            _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, phases & first, q, slot, st))
            if phases & _capi.PHASE_BEFORE:
                self._wrap.run(field)
        rest = phases & ~first
        if rest:
            _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, rest, q, slot, st))

    def _p2p_refresh(self):
        """push both boundary planes and wait for the neighbours' (collective; after the user wrote E / H)."""
        _capi.check(self.lib, self.lib.fdtd_halo_refresh(C.byref(self.desc), C.byref(self._p2p.h), self._stream()))

    def _p2p_halfstep(self, field, q, slot):
        """one half-step of an x-sharded slab with peer-to-peer ghost planes: one C call (fdtd_sharded_halfstep)."""
        lib, d, h = self.lib, self.desc, self._p2p.h
        fidx = 0 if field == "E" else 1
        if self._wrap is None:
            _capi.check(lib, lib.fdtd_sharded_halfstep(C.byref(d), C.byref(h), fidx, q, slot, self._stream()))
            return
        # a periodic x boundary crosses the slabs: the wrap plane travels between the post ops registered before and
        # after it (send / recv of the process group), so the parts are driven from here
        n = d.Nx
        step = lib.fdtd_e_halfstep if field == "E" else lib.fdtd_h_halfstep
        bulk = (1, n) if field == "E" else (0, n - 1)
        edge = (0, min(1, n)) if field == "E" else (max(n - 1, 0), n)
        st = self._stream()
        main, side = torch.cuda.current_stream(self.grid._E.device), self._p2p.stream
        sst = C.c_void_p(side.cuda_stream)
        side.wait_stream(main)
        _capi.check(lib, step(C.byref(d), bulk[0], bulk[1], q, slot, st))
        has_nb = bool(h.has_left if field == "E" else h.has_right)
        if has_nb:
            _capi.check(lib, lib.fdtd_halo_wait(C.c_void_p(h.flags + 8 * (1 - fidx)), h.count[1 - fidx],
                                                C.c_void_p(h.error), h.timeout_ns, sst))
        _capi.check(lib, step(C.byref(d), edge[0], edge[1], q, slot, sst))
        main.wait_stream(side)
        self._post(field, q, slot, st)
        if has_nb:
            gy, gz, flag = ((h.left_ghost_y, h.left_ghost_z, h.left_flag) if field == "E"
                            else (h.right_ghost_y, h.right_ghost_z, h.right_flag))
            side.wait_stream(main)
            _capi.check(lib, lib.fdtd_halo_push(C.byref(d), fidx, C.c_void_p(gy), C.c_void_p(gz), sst))
            _capi.check(lib, lib.fdtd_halo_signal(C.c_void_p(flag), h.count[fidx] + 1, sst))
        h.count[fidx] += 1

    def _classify(self, ie_eff, ie_grid_if_objects, absorb, imu, ty, tz, ie2=None, absorb2=None):
        """per-(plane, y-tile, z-tile) class byte, FDTD_CLS_* (include/fdtd_b200.h)."""
        g = self.grid
        nx, Ny, Nz = g._part.nx, g.Ny, g.Nz
        nty, ntz = -(-Ny // ty), -(-Nz // tz)
        cls = torch.zeros((nx, nty, ntz), dtype=torch.uint8, device=g._E.device)
        dt = g._dtype

        def tiles(mask):        # (n, Ny, Nz) bool -> (n, nty, ntz) bool
            m = torch.nn.functional.pad(mask, (0, ntz * tz - Nz, 0, nty * ty - Ny))
            return m.view(m.shape[0], nty, ty, ntz, tz).any(dim=4).any(dim=2)

        bg_e = torch.tensor(g._bg_inv_eps, dtype=dt, device=cls.device).view(3, 1, 1, 1)
        bg_m = torch.tensor(g._bg_inv_mu, dtype=dt, device=cls.device).view(3, 1, 1, 1)
        step = max(1, (64 << 20) // max(1, Ny * Nz))      # bound the temporaries
        for a in range(0, nx, step):
            b = min(nx, a + step)
            bits = torch.zeros((b - a, nty, ntz), dtype=torch.uint8, device=cls.device)
            if ie_eff is not None:
                bits |= tiles((ie_eff[:, a:b] != bg_e).any(0)).to(torch.uint8) * _capi.CLS_VARY_E
            if imu is not None:
                bits |= tiles((imu[:, a:b] != bg_m).any(0)).to(torch.uint8) * _capi.CLS_VARY_H
            if absorb is not None:
                bits |= tiles((absorb[:, a:b] != 0).any(0)).to(torch.uint8) * _capi.CLS_ABSORB
            if ie2 is not None:
                bits |= tiles((ie2[:, a:b] != 0).any(0)).to(torch.uint8) * _capi.CLS_OVERLAP
            if absorb2 is not None:
                bits |= tiles((absorb2[:, a:b] != 0).any(0)).to(torch.uint8) * _capi.CLS_ABSORB2
            if ie_grid_if_objects is not None:
                bits |= tiles(((ie_eff[:, a:b] != ie_grid_if_objects[:, a:b])
                               | (False if ie2 is None else ie2[:, a:b] != 0)).any(0)).to(torch.uint8) * _capi.CLS_OBJECT
                gx = ie_grid_if_objects[0:2, a:b]
                bits |= tiles(((gx == 0) & torch.signbit(gx)).any(0)).to(torch.uint8) * _capi.CLS_ANISO
            cls[a:b] = bits
        return cls.contiguous()

    def _source_signature(self):
        """the per-step parameters the reference re-reads from every source on every step (fdtd/sources.py:95-108,
        280-295, 478-486, 601-626): changing one between steps must take effect on the next step."""
        return tuple(s._signature() for s in self.grid.sources if hasattr(s, "_signature"))

    def stale(self):
        g = self.grid
        if g._baked_counts != g._registration_count:
            return True
        if self._src_sig != self._source_signature():
            return True
        v = (None if g._inv_eps is None else g._inv_eps._version,
             None if g._inv_mu is None else g._inv_mu._version)
        return v != self._mat_versions

    # ------------------------------------------------------------------------------ per-run state
    def _ensure_wave(self, q0, n, lookahead=WAVE_TABLE_MIN):
        """host-tabulated per-step source scalars covering steps [q0, q0+n); `lookahead`: how far beyond to tabulate
        (step()-driven loops ask for one step at a time; run() knows its chunk)."""
        if not self._src_entries:
            return
        if self._wave is not None:
            a, ln = self._wave
            if a <= q0 and q0 + n <= a + ln:
                return
        ln = max(n, lookahead)
        g, d = self.grid, self.desc
        tables = {}
        for idx, src in self._src_entries:
            if id(src) not in tables:
                vals = torch.tensor([src._wave_value(q) for q in range(q0, q0 + ln)], dtype=torch.float64)
                tables[id(src)] = vals.to(g._dtype).to(g._E.device)
            t = tables[id(src)]
            e = d.sources[idx]
            e.wave, e.wave_q0, e.wave_len = _ptr(t), q0, ln
            if e.kind == _capi.SRC_FEEDBACK:
                key = ("div", id(src))
                if key not in tables:
                    vals = torch.tensor([src._wave_value(q) / g.grid_spacing for q in range(q0, q0 + ln)],
                                        dtype=torch.float64)
                    tables[key] = vals.to(g._dtype).to(g._E.device)
                e.profile = _ptr(tables[key])
        self._wave_keep = list(tables.values())
        self._wave = (q0, ln)

    def _stream(self):
        dev = self.grid._E.device
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else C.c_void_p(None)

    def _slot(self, field):
        """next free ring slot for `field`, flushing the rings to the host when full."""
        g = self.grid
        if not self._dets and not self._feedback_sources:
            return 0
        if g._ring_fill[field] >= self.ring_capacity:
            self.flush_detectors()
        return g._ring_fill[field]

    def flush_detectors(self):
        g = self.grid
        nE, nH = g._ring_fill["E"], g._ring_fill["H"]
        if nE == 0 and nH == 0:
            return
        for det in self._dets:
            det._drain(nE, nH)
        for _, src in self._feedback_sources:
            src._drain(nE)
        g._ring_fill["E"] = g._ring_fill["H"] = 0
        if self._p2p:
            self._p2p.check()        # the drain synchronised anyway: a timed-out halo wait surfaces here at the latest

    # ------------------------------------------------------------------------------------ stepping
    def _sharded_halfstep(self, field, q, slot):
        """one half-step of an x-sharded slab.  The plane that needs the neighbour's ghost plane (plane 0
        for E, the last plane for H) runs on the halo stream right behind the exchange that delivers the
        ghost, concurrently with the bulk on the main stream; the exchange of the freshly updated
        boundary plane is then started and overlaps the bulk of the NEXT half-step."""
        if self._p2p:
            return self._p2p_halfstep(field, q, slot)
        lib, d, halo = self.lib, self.desc, self._halo
        n = d.Nx
        step = lib.fdtd_e_halfstep if field == "E" else lib.fdtd_h_halfstep
        bulk = (1, n) if field == "E" else (0, n - 1)
        edge = (0, min(1, n)) if field == "E" else (max(n - 1, 0), n)
        other = "H" if field == "E" else "E"
        st = self._stream()
        if halo.cuda:
            dev = self.grid._E.device
            main, side = torch.cuda.current_stream(dev), halo.stream
            side.wait_stream(main)                       # everything enqueued so far (user writes included)
            _capi.check(lib, step(C.byref(d), bulk[0], bulk[1], q, slot, st))
            _capi.check(lib, step(C.byref(d), edge[0], edge[1], q, slot, C.c_void_p(side.cuda_stream)))
            main.wait_stream(side)
        else:
            _capi.check(lib, step(C.byref(d), bulk[0], bulk[1], q, slot, st))
            halo.wait(self._pending[other])
            _capi.check(lib, step(C.byref(d), edge[0], edge[1], q, slot, st))
        self._pending[other] = None
        if field == "H" and self._late:
            # a CurrentDetector on the first plane of a slab reads the neighbour's H of THIS half-step, in its final
            # state: post ops and sources first, then the exchange, then the detectors
            self._post(field, q, slot, st, _capi.PHASE_ALL & ~_capi.PHASE_DETECTORS)
            halo.wait(halo.start(field))
            self._post(field, q, slot, st, _capi.PHASE_DETECTORS)
            self._pending[field] = None
            return
        self._post(field, q, slot, st)
        self._pending[field] = halo.start(field)

    def _hooked_halfstep(self, field, q, slot):
        """one half-step with user plug-ins in the loop (unsharded), in the reference's order (fdtd/grid.py:275-325):
        boundaries' update_phi -> field update + built-in objects (kernel) -> deeper built-in objects -> user objects'
        update(curl) -> boundaries' update (built-in post ops, then the user's) -> sources (built-in, then the user's) ->
        detectors (built-in, then the user's).  Within a phase the built-in plug-ins run before the user's."""
        g, lib, d, hk = self.grid, self.lib, self.desc, self._hooks
        fidx = 0 if field == "E" else 1
        st = self._stream()

        def call(kind, method, *args):
            for thing, methods in hk[kind]:
                if method in methods:
                    getattr(thing, method)(*args)

        call("boundaries", "update_phi_" + field)
        step = lib.fdtd_e_halfstep if field == "E" else lib.fdtd_h_halfstep
        curl = None
        if any("update_" + field in m for _, m in hk["objects"]):
            from .grid import curl_E, curl_H        # the curl the update is about to use, for the user's objects
            curl = curl_H(g.H) if field == "E" else curl_E(g.E)
        _capi.check(lib, step(C.byref(d), 0, d.Nx, q, slot, st))
        _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, _capi.PHASE_OBJECTS, q, slot, st))
        call("objects", "update_" + field, curl)
        _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, _capi.PHASE_BEFORE | _capi.PHASE_AFTER, q, slot, st))
        call("boundaries", "update_" + field)
        _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, _capi.PHASE_SOURCES, q, slot, st))
        call("sources", "update_" + field)
        _capi.check(lib, lib.fdtd_post_phases(C.byref(d), fidx, _capi.PHASE_DETECTORS, q, slot, st))
        call("detectors", "detect_" + field)

    def update_E(self, q):
        g, lib, d = self.grid, self.lib, self.desc
        self._ensure_wave(q, 1)
        slot = self._slot("E")
        if self._hooked:
            self._hooked_halfstep("E", q, slot)
        elif self._halo is None:
            _capi.check(lib, lib.fdtd_update_E(C.byref(d), q, slot, self._stream()))
        else:
            self._sharded_halfstep("E", q, slot)
        if self._dets or self._feedback_sources:
            g._ring_fill["E"] += 1
        for _, src in self._feedback_sources:
            src._steps_logged.append(q)

    def update_H(self, q):
        g, lib, d = self.grid, self.lib, self.desc
        self._ensure_wave(q, 1)
        slot = self._slot("H")
        if self._hooked:
            self._hooked_halfstep("H", q, slot)
        elif self._halo is None:
            _capi.check(lib, lib.fdtd_update_H(C.byref(d), q, slot, self._stream()))
        else:
            self._sharded_halfstep("H", q, slot)
        if self._dets or self._feedback_sources:
            g._ring_fill["H"] += 1

    def run(self, q0, nsteps, progress=None):
        """nsteps full steps from step index q0; one C call per chunk when not sharded."""
        g, lib, d = self.grid, self.lib, self.desc
        done = 0
        rings = bool(self._dets or self._feedback_sources)
        while done < nsteps:
            if rings and (g._ring_fill["E"] != g._ring_fill["H"] or g._ring_fill["E"] >= self.ring_capacity):
                self.flush_detectors()
            room = self.ring_capacity - g._ring_fill["E"] if rings else nsteps
            n = min(nsteps - done, room, WAVE_TABLE_MAX)
            q = q0 + done
            self._ensure_wave(q, n, lookahead=min(WAVE_TABLE_MIN, -(-n // 32) * 32))
            if self._hooked:
                # user plug-ins read grid.time_steps_passed like the reference's sources do: keep it current per step
                for s in range(n):
                    g.time_steps_passed = q + s
                    self.update_E(q + s)
                    self.update_H(q + s)
            elif self._halo is None:
                _capi.check(lib, lib.fdtd_run(C.byref(d), q, n, g._ring_fill["E"] if rings else 0,
                                              self._stream()))
                if rings:
                    g._ring_fill["E"] += n
                    g._ring_fill["H"] += n
                for _, src in self._feedback_sources:
                    src._steps_logged.extend(range(q, q + n))
            elif self._p2p and self._wrap is None:
                # the whole chunk in one C call: the ranks only meet through the flag words
                _capi.check(lib, lib.fdtd_run_sharded(C.byref(d), C.byref(self._p2p.h), q, n,
                                                      g._ring_fill["E"] if rings else 0, self._stream()))
                if rings:
                    g._ring_fill["E"] += n
                    g._ring_fill["H"] += n
                for _, src in self._feedback_sources:
                    src._steps_logged.extend(range(q, q + n))
            else:
                for s in range(n):
                    self.update_E(q + s)
                    self.update_H(q + s)
            done += n
            g.time_steps_passed = q0 + done      # per chunk: an interrupted run leaves counter and fields in step
            if progress is not None:
                progress.update(n)

    def quiesce(self):
        """make every enqueued halo exchange visible to the current stream (before field reads)."""
        if self._halo is not None:
            if not self._p2p:
                for f in ("E", "H"):
                    self._halo.wait(self._pending[f])
                    self._pending[f] = None
            if self._halo.cuda:
                torch.cuda.current_stream(self.grid._E.device).wait_stream(self._halo.stream)
