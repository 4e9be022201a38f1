"""x-slab sharding over NCCL on real GPUs: sharded == single-GPU result, bit for bit (needs >= 2 GPUs)."""
import numpy as np
import pytest
import torch

import scenes
from test_sharded_gloo import launch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("scene,dtype", [("pml3d", "float64"), ("objects3d", "float32"), ("c4small", "float32"),
                                         ("periodic3d", "float64"), ("ring3d", "float32"),
                                         ("overlaps3d", "float64")])
def test_sharded_equals_single(tmp_path, scene, dtype, halo):
    """halo = p2p: ghost planes stored straight into the neighbour's memory (CUDA IPC peer pointers + flags);
    halo = nccl: send/recv.  Both must reproduce the single-GPU run bit for bit."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, torch.cuda.device_count())
    steps = 30
    out = str(tmp_path / "sharded.npz")
    launch(world, "nccl", dtype, scene, steps, out, FDTD_B200_HALO=halo)
    got = dict(np.load(out))
    import fdtd_b200 as fd
    fd.set_backend("cuda." + dtype)
    g = scenes.SCENES[scene][0](fd)
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


@pytest.mark.parametrize("scene,dtype,steps,chunk", [("fusedslab", "float32", 31, 0), ("fusedslab", "float64", 26, 3),
                                                     ("fusedslab", "float32", 24, 2), ("fusedslab", "float32", 27, 4)])
def test_temporally_fused_steps_on_slabs(tmp_path, scene, dtype, steps, chunk):
    """x-sharded grids with grid._fuse_eh = 1: every rank runs pairs of single-pass E+H steps (the fused kernel stores
    E_new[plane 0] into the left neighbour's second buffer, the last H plane follows once the right neighbour's E_new
    has arrived), an odd remainder and step()-driven steps run as two half-steps.  Bit-identical to the single-GPU
    two-half-step run."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, torch.cuda.device_count())
    out = str(tmp_path / "sharded.npz")
    # chunk > 0: the slab is cut into several x-chunks -- the first and the last one run on the side stream with the
    # flags and the last H plane, the ones in between on the main stream
    launch(world, "nccl", dtype, scene, steps, out, FDTD_TEST_FUSE_EH="1", FDTD_TEST_X_CHUNK=str(chunk))
    got = dict(np.load(out))
    import fdtd_b200 as fd
    fd.set_backend("cuda." + dtype)
    build = scenes.SCENES[scene][0] if scene in scenes.SCENES else getattr(scenes, scene)
    g = build(fd)
    g._fuse_eh = 0
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    assert float(np.abs(want["E"]).max()) > 0
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


def test_energy_slices_gather_one_plane(tmp_path):
    """`energy_slice` on an x-sharded grid on real GPUs: the x-plane comes from its owner, y / z planes are gathered
    slab by slab -- equal to the energy of the gathered fields."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(4, torch.cuda.device_count())
    out = str(tmp_path / "sharded.npz")
    launch(world, "nccl", "float64", "objects3d", 20, out, FDTD_TEST_SLICES="7,9,4")
    got = dict(np.load(out))
    energy = (got["E"] ** 2 + got["H"] ** 2).sum(-1)
    assert np.array_equal(got["slice_x"], energy[7])
    assert np.array_equal(got["slice_y"], energy[:, 9, :].T)
    assert np.array_equal(got["slice_z"], energy[:, :, 4])
