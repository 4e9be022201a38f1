"""worker of tests/test_sharded_*.py: one rank of a torch.distributed job running a scene x-sharded.

    RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT in the env;  argv: backend(gloo|nccl) dtype scene steps out.npz
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    backend, dtype, scene, steps, out = sys.argv[1:6]
    steps = int(steps)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if backend == "gloo":
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", torch.cuda.current_device()))
    if "@" in scene:
        # several "scene@dtype" items in one job (saves the start-up of the ranks): out -> out.<n>.npz
        for n, item in enumerate(scene.split(",")):
            name, dt = item.split("@")
            run_one(backend, dt, name, steps, f"{out}.{n}.npz", rank, world)
    else:
        run_one(backend, dtype, scene, steps, out, rank, world)
    if backend != "gloo":
        torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def run_one(backend, dtype, scene, steps, out, rank, world):
    import scenes
    if backend == "gloo":
        from emu.harness import use_emu
        fd = use_emu(dtype)
    else:
        import fdtd_b200 as fd
        fd.set_backend("cuda." + dtype)
    if scene.startswith("fuzz:"):
        from fuzz_scenes import random_scene
        build, _ = random_scene(int(scene[5:]))
    else:
        build = scenes.SCENES[scene][0] if scene in scenes.SCENES else getattr(scenes, scene)
    if os.environ.get("FDTD_TEST_RING_BYTES"):           # tiny detector rings: several collective flushes per run
        import fdtd_b200.engine as engine
        engine.RING_BYTES = int(os.environ["FDTD_TEST_RING_BYTES"])
    try:
        g = build(fd)
        if os.environ.get("FDTD_TEST_FUSE_EH"):
            g._fuse_eh = int(os.environ["FDTD_TEST_FUSE_EH"])
        if os.environ.get("FDTD_TEST_X_CHUNK"):      # short x-chunks: a slab of a small scene is cut into several
            g._x_chunk = int(os.environ["FDTD_TEST_X_CHUNK"])
        if os.environ.get("FDTD_TEST_TRACK"):
            scenes.track_all(g, steps)
        g.run(0, progress_bar=False)           # bake: sharding restrictions surface here, on every rank alike
    except (NotImplementedError, ValueError) as exc:
        if rank == 0:
            np.savez(out, skipped=np.array(str(exc)))
        dist.barrier()
        return
    assert g._part.world == world and g._part.sharded
    if os.environ.get("FDTD_TEST_EXPECT_LATE"):      # a CurrentDetector on a slab's first plane: exchange-then-sample path
        assert g._engine._late and not g._engine._p2p
    half = steps // 2
    g.run(half, progress_bar=False)
    for _ in range(steps - half):          # exercise the step()-granular path too
        g.step()
    res = scenes.dump(g)                   # grid.E gathers the slabs, detectors gather their samples
    if os.environ.get("FDTD_TEST_TRACK"):
        res.update(scenes.dump_tracked(g))
    if os.environ.get("FDTD_TEST_SLICES"):
        from fdtd_b200.visualization import energy_slice
        ix, iy, iz = (int(v) for v in os.environ["FDTD_TEST_SLICES"].split(","))
        res.update(slice_x=energy_slice(g, x=ix), slice_y=energy_slice(g, y=iy), slice_z=energy_slice(g, z=iz))
    if rank == 0:
        np.savez(out, **res)
    if os.environ.get("FDTD_TEST_FUSE_EH") == "1":
        import ctypes
        assert g._engine.lib.fdtd_fuse_eh_sharded_active(ctypes.byref(g._engine.desc), ctypes.byref(g._engine._p2p.h)) == 1
    if backend != "gloo":
        if os.environ.get("FDTD_B200_HALO", "p2p") == "p2p":
            assert g._engine._p2p, "peer-to-peer halo was requested but not set up"
        torch.cuda.synchronize()
    del g
    dist.barrier()


if __name__ == "__main__":
    main()
