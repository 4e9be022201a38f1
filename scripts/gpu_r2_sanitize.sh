#!/bin/bash
# compute-sanitizer on the kernels (development tool; logs are copied to profiles/): memcheck, racecheck, synccheck,
# initcheck on small scenes of every kernel family incl. the fused E+H kernel, then memcheck over a 2-rank
# peer-to-peer run when two GPUs are present.
set -u
out=gpurun_out/r2_sanitize
mkdir -p $out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_scenes.py fused > $out/$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|error" $out/$tool.log | tail -5
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  cat > /tmp/p2p_scene.py <<'PY'
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests")
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
import fdtd_b200 as fd, scenes
fd.set_backend("cuda.float32")
g = scenes.c4small(fd)
g.run(20, progress_bar=False)
for _ in range(4):
    g.step()
out = scenes.dump(g)
torch.cuda.synchronize()
assert g._engine._p2p
if dist.get_rank() == 0:
    print("p2p run done", float(np.abs(out["E"]).max()), flush=True)
dist.barrier(); dist.destroy_process_group()
PY
  timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 /tmp/p2p_scene.py > $out/memcheck_p2p_2gpu.log 2>&1
  echo "== memcheck p2p: rc=$?"; grep -E "ERROR SUMMARY|p2p run done|Error" $out/memcheck_p2p_2gpu.log | tail -6
fi
ls -la $out
