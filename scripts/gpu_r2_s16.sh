#!/bin/bash
# Round-2 GPU session 16 (development tool, 1 GPU): what an x-PML plane costs in the fused kernel and in the two
# half-steps (the weights of the cost-balanced partition), and the L2 promotion of the tensor maps.
set -u
out=gpurun_out/r2_s16
mkdir -p $out
python - <<'PY' 2>&1 | tee $out/xpml_cost.log
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fdtd_b200 as fd
fd.set_backend("cuda.float32")
def build(nx, tx, t=10, n=1024):
    g = fd.Grid(shape=(nx, n, n), grid_spacing=77.5e-9)
    if tx:
        g[0:tx, :, :] = fd.PML(); g[-tx:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML(); g[:, -t:, :] = fd.PML(); g[:, :, 0:t] = fd.PML(); g[:, :, -t:] = fd.PML()
    g[nx // 2, n // 2, n // 2] = fd.PointSource(period=20)
    return g
def t(nx, tx, fuse):
    g = build(nx, tx); g._fuse_eh = fuse
    g.run(4, progress_bar=False); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(40, progress_bar=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    del g
    return ms
for fuse in (1, 0):
    r = {tx: t(192, tx, fuse) for tx in (0, 10, 30)}
    big = t(384, 10, fuse)
    plane = (big - r[10]) / 192
    pml = (r[30] - r[10]) / 40
    print(f"fuse_eh={fuse}: 192 planes with x-PML 0/10/30 cells per face: {r[0]:.4f} / {r[10]:.4f} / {r[30]:.4f} ms; 384 planes: {big:.4f} ms"
          f" -> ordinary plane {plane * 1e3:.2f} us, x-PML plane {plane * 1e3 + pml * 1e3:.2f} us (x {1 + pml / plane:.2f})", flush=True)
PY
{
for p in 2 0 1 3; do
  echo "# FDTD_B200_TMA_L2PROMO=$p"; FDTD_B200_TMA_L2PROMO=$p python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
} | tee $out/l2promo.log
