"""GPU tests of the x-slab path: the halo exchange is inside the stencil kernel (peer stores + flags).
* several slabs on ONE device (peer pointer = local pointer, shared stream): runs on the 1-GPU box and
  exercises the whole protocol -- boundary stores, CTA counters, flags, ghost ownership of sources;
* several devices driven by one process, and one process per GPU over CUDA IPC + torch.distributed (NCCL):
  skipped unless the box has >= 2 GPUs.
The result must be BIT-IDENTICAL to the single-slab run and to the oracle."""
import os
import socket

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT, bits_equal

pytestmark = pytest.mark.gpu


def _case(seed, shape, T, S, seam_parts=0):
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
    m = rng.uniform(0.5, 3.0, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    crd = (rng.uniform(-0.04, 1.04, (S, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    if seam_parts:
        base, i = nx // seam_parts, 0
        for k in range(1, seam_parts):  # sources straddling every seam
            crd[i % S, 0] = np.float32((k * base - 1) * 0.1) + np.float32(0.04)
            crd[(i + 1) % S, 0] = np.float32(k * base * 0.1)
            i += 2
    return u, m, src, crd


@pytest.mark.parametrize("nparts,shape,opts", [
    (2, (24, 16, 64), {}),
    (3, (40, 20, 72), {"exact": 0}),
    (2, (64, 24, 128), {"tile_y": 8, "tile_z": 128, "rows": 2, "xchunk": 8}),
    (4, (37, 16, 64), {"xchunk": 5}),
])
def test_slabs_on_one_device_match_single_slab_and_oracle(pkg, oracle, nparts, shape, opts):
    T, S = 9, 8
    u, m, src, crd = _case(31 + nparts, shape, T, S, seam_parts=nparts)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    # single slab
    one = u.copy()
    with pkg.Plan(*shape, deviceid=0) as p:
        for k, v in opts.items():
            p.set_option(k, v)
        p.upload(one, m)
        p.set_sources(src, crd)
        p.run(0, T - 1)
        p.download(one)
    # nparts slabs on the same device
    ls = pkg.LocalSlabs(shape[0], shape[1], shape[2], [0] * nparts, options=opts)
    ls.upload(u, m)
    ls.set_sources(src, crd)
    t = ls.run(0, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, one), "slab run differs from the single-slab run"
    if opts.get("exact", 1):
        assert bits_equal(out, ref)
    else:
        assert oracle.rel_l2(out, ref) < 1e-4
    assert t.section0 > 0


def test_slab_run_split_in_two_calls(pkg, oracle):
    """Ghost planes stay valid across runs (restart with time_m > 0)."""
    shape, T, S = (32, 16, 64), 14, 4
    u, m, src, crd = _case(9, shape, T, S, seam_parts=2)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    ls = pkg.LocalSlabs(*shape, [0, 0])
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, 5)
    ls.run(6, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, ref)


def test_slabs_on_two_devices_one_process(pkg, oracle):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    shape, T, S = (96, 64, 128), 12, 6
    u, m, src, crd = _case(5, shape, T, S, seam_parts=2)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", threads=8)
    ls = pkg.LocalSlabs(*shape, [0, 1])
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, shape, T, S, out_path, t_fuse=1):
    import importlib
    import sys

    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    pkg = importlib.import_module(PKG_NAME)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    if t_fuse == 2:
        from test_tb2_gpu import fused_case
        u, m, src, crd = fused_case(5, shape, T, S, seam_parts=world)
    else:
        u, m, src, crd = _case(5, shape, T, S, seam_parts=world)
    sr = pkg.SlabRun(dist, shape[0], shape[1], shape[2], rank)
    sr.plan.set_option("t_fuse", t_fuse)
    sr.plan.upload(pkg.slab.slab_view(u, sr.x_offset, sr.nx), pkg.slab.slab_view(m, sr.x_offset, sr.nx))
    sr.plan.set_sources(src, crd)
    dist.barrier()
    sr.run(0, T // 2)
    sr.run(T // 2 + 1, T - 1)
    assert sr.plan.get_option("t_fuse_used") == t_fuse
    mine = sr.plan.download()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out = np.zeros_like(u)
        pkg.slab.assemble(out, gathered, sr.parts)
        np.save(out_path, out)
    sr.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("t_fuse", [1, 2])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu_over_ipc(pkg, oracle, tmp_path, world, t_fuse):
    """t_fuse = 2: two-step passes, 4-plane ghost zones, the depth negotiated over torch.distributed."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    shape, T, S = (32 * world, 64, 128), 13, 6
    if t_fuse == 2:
        from test_tb2_gpu import fused_case
        u, m, src, crd = fused_case(5, shape, T, S, seam_parts=world)
    else:
        u, m, src, crd = _case(5, shape, T, S, seam_parts=world)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", threads=8)
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_ipc_worker, args=(world, _free_port(), shape, T, S, out_path, t_fuse), nprocs=world, join=True)
    assert bits_equal(np.load(out_path), ref)


@pytest.mark.parametrize("t_fuse", [1, 2])
def test_linked_slabs_fuse_sources_even_when_fusion_is_off(pkg, oracle, t_fuse):
    """fuse_inject = 0 asks for the stand-alone scatter, but a linked slab pushes its boundary planes to the neighbour
    inside the Section0 launch: a source on a boundary plane must already be in them.  Linked slabs therefore always
    fuse their interior cells; sources sit on and next to the seam."""
    from oracle import windows as W

    shape, T, S = (64, 32, 64), 11, 8
    u, m, src, crd = W.dense_seam_case(5, shape, T, S, 2)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    ls = pkg.LocalSlabs(*shape, [0, 0], options={"fuse_inject": 0, "t_fuse": t_fuse})
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, ref)


def test_pairing_does_not_depend_on_source_ownership(pkg, oracle):
    """The src table ends before the run does (src_size0 - 1 < time_M): the step that crosses the end of src must not be
    paired, on EVERY slab -- including the one that owns no source cell (the decision used to look at slab 0)."""
    from oracle import windows as W

    shape, T = (96, 32, 64), 16
    u, m, src, crd = W.dense_seam_case(6, shape, T, 3, 1)
    crd[:, 0] = np.float32(80 * 0.1) + np.float32(0.03)  # every source in the LAST slab; slab 0 owns none
    src = src[:10]                                        # src ends at step 9 (odd offset from the timed boundary)
    ref = u.copy()
    # the reference reads src[time] for every step (openacc.cpp:134): give the oracle zero rows past the end, which is
    # what "no injection" means bit for bit on a field without negative zeros
    oracle.run(ref, m, np.concatenate([src, np.zeros((T - 10, src.shape[1]), np.float32)]), crd, impl="port", time_M=T - 1)
    ls = pkg.LocalSlabs(*shape, [0, 0, 0], options={"t_fuse": 2})
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    assert ls.plans[0].get_option("t_fuse_used") == 2
    ls.close()
    assert bits_equal(out, ref)
