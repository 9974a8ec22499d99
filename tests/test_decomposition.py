"""Slab decomposition of one body (BASELINE config 5): host partition logic on CPU (incl. a world-size-2
gloo run of the face-list agreement) and, on the GPU, bit-exact agreement of a body stepped as two / three slabs with halo
exchange against the same body stepped whole."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from util import cube_spec

import __graft_entry__ as graft

graft.load_package()
from voxcraft_sim_b200 import abi, parallel  # noqa: E402


def build_full(n=(9, 4, 3), seed=5, holes=0.0):
    spec = cube_spec(n, seed=seed, actuated=True, holes=holes, name="decomp")
    lib = util.load_engine()
    b, d = spec.build(lib)
    return lib, b, d


def check_partition(d, world):
    full = d.contents
    nv, nl = full.n_voxels, full.n_links
    slabs = [parallel.partition_slabs(d, world, r) for r in range(world)]
    owned_all = np.concatenate([s.voxels[s.owned] for s in slabs])
    assert sorted(owned_all.tolist()) == list(range(nv)), "every voxel is owned by exactly one rank"
    link_seen = np.zeros(nl, np.int32)
    vneg = np.ctypeslib.as_array(full.link_vneg, shape=(nl,))
    vpos = np.ctypeslib.as_array(full.link_vpos, shape=(nl,))
    for s in slabs:
        sd = s.desc
        n_v, n_l = sd.n_voxels, sd.n_links
        ln = np.ctypeslib.as_array(sd.link_vneg, shape=(n_l,))
        lp = np.ctypeslib.as_array(sd.link_vpos, shape=(n_l,))
        la = np.ctypeslib.as_array(sd.link_axis, shape=(n_l,))
        vl = np.ctypeslib.as_array(sd.vox_links, shape=(n_v * 6,)).reshape(n_v, 6)
        fl = np.ctypeslib.as_array(sd.vox_flags, shape=(n_v,))
        assert ((fl & abi.VOX_GHOST) != 0).tolist() == (~s.owned).tolist()
        # owned voxels first, ghosts last (the engine's voxel pass runs over the owned prefix), each group in the body's own order;
        # links with two owned ends first, links across a face last (the link pass meets its ghosts only in the last tiles)
        n_own = int(s.owned.sum())
        assert s.owned[:n_own].all() and not s.owned[n_own:].any()
        assert (np.diff(s.voxels[:n_own]) > 0).all() and (np.diff(s.voxels[n_own:]) > 0).all()
        face = ~(s.owned[ln] & s.owned[lp])
        assert not face[: int((~face).sum())].any()
        assert (s.owned[ln] | s.owned[lp]).all(), "every link of a slab has an owned end"
        # the engine's slot rule (2*axis at the negative end, 2*axis+1 at the positive end) survives the renumbering
        assert (vl[ln, 2 * la] == np.arange(n_l)).all() and (vl[lp, 2 * la + 1] == np.arange(n_l)).all()
        # an owned voxel keeps all its links; the global link behind each local one has the same ends
        gl = {(int(s.voxels[a]), int(s.voxels[b])) for a, b in zip(ln, lp)}
        for g in np.nonzero(s.owned)[0]:
            gv = int(s.voxels[g])
            want = {(int(a), int(b)) for a, b in zip(vneg, vpos) if a == gv or b == gv}
            assert want <= gl
        for a, b in gl:
            link_seen[np.nonzero((vneg == a) & (vpos == b))[0]] += 1
    assert (link_seen >= 1).all() and (link_seen <= 2).all()
    for r in range(world - 1):  # what r sends up is what r+1 receives from below, in the same order (and vice versa)
        lo, hi = slabs[r], slabs[r + 1]
        assert lo.voxels[lo.send[1]].tolist() == hi.voxels[hi.recv[0]].tolist()
        assert hi.voxels[hi.send[0]].tolist() == lo.voxels[lo.recv[1]].tolist()
        assert lo.owned[lo.send[1]].all() and not hi.owned[hi.recv[0]].any()
    return slabs


def test_partition_invariants():
    for n, holes, world in (((9, 4, 3), 0.0, 2), ((12, 3, 3), 0.25, 3), ((8, 5, 2), 0.1, 4)):
        lib, b, d = build_full(n, seed=11, holes=holes)
        try:
            check_partition(d, world)
        finally:
            lib.vx3_builder_destroy(b)


def test_slab_bounds_balance():
    coord = np.repeat(np.arange(10), [5, 1, 1, 1, 1, 1, 1, 1, 1, 5])
    b = parallel.slab_bounds(coord, 2)
    assert b[0] == 0 and b[-1] == 10 and 0 < b[1] < 10
    with pytest.raises(ValueError):
        parallel.slab_bounds(np.array([0, 0, 1]), 3)


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib, b, d = build_full((10, 3, 3), seed=3, holes=0.15)
    s = parallel.partition_slabs(d, world, rank)
    mine = {side: (s.voxels[s.send[side]].tolist(), s.voxels[s.recv[side]].tolist()) for side in (0, 1)}
    table = [None] * world
    dist.all_gather_object(table, mine)
    ok = True
    for side, nb in ((0, rank - 1), (1, rank + 1)):
        if 0 <= nb < world:
            ok &= table[nb][1 - side][1] == mine[side][0] and table[nb][1 - side][0] == mine[side][1]
    q.put((rank, ok, int(s.owned.sum())))
    dist.barrier()
    dist.destroy_process_group()
    lib.vx3_builder_destroy(b)


def test_face_lists_agree_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got)
    lib, b, d = build_full((10, 3, 3), seed=3, holes=0.15)
    assert sum(n for _, _, n in got) == d.contents.n_voxels
    lib.vx3_builder_destroy(b)


@pytest.mark.gpu
@pytest.mark.parametrize("world,exchange", [(2, "kernels"), (3, "kernels"), (2, "in_step"), (3, "in_step")])
def test_slabs_with_halo_exchange_match_whole_body_bit_exactly(world, exchange, monkeypatch):
    """One process drives all slabs (batches on one device, wired with halo_connect_local) against the whole body: the face
    links are evaluated on both sides from identical inputs, so every owned voxel must agree bit for bit.  `kernels`: the
    stand-alone send / wait / receive kernels (the default between slabs of one process); `in_step`: the multi-GPU default —
    the voxel pass sends, the link pass reads the receive buffers (vx3_halo.cuh) — forced onto the one-process set-up."""
    from util import EngineBatch
    if exchange == "in_step":
        monkeypatch.setenv("VX3_HALO_INKERNEL", "2")
        monkeypatch.setenv("VX3_HALO_TIMEOUT_MS", "5000")
    lib, b, d = build_full((9, 4, 3), seed=21, holes=0.1)
    try:
        dt = float(np.float32(0.9 * lib.vx3_model_recommended_dt(d)))
        whole = EngineBatch([d])
        whole.set_profiling(False, use_persistent=False)
        slabs = [parallel.partition_slabs(d, world, r) for r in range(world)]
        parts = [parallel.DecomposedBody(s, dt) for s in slabs]
        for r in range(world):
            for side, nb in ((0, r - 1), (1, r + 1)):
                if 0 <= nb < world:
                    parts[r].batch.halo_connect_local(side, parts[nb].batch)
        for chunk in range(4):
            whole.step(150, dt)
            for _ in range(5):  # one host thread feeds all slabs: keep each round far below the driver's launch-queue depth, or
                for p in parts:  # the host blocks queueing one slab while that slab's stream waits for a slab not yet queued
                    p.batch.step_async(30, dt)
                for p in parts:
                    p.batch.sync()
            sw = whole.state(0)
            for s, p in zip(slabs, parts):
                sp = p.batch.state(0)
                own = np.nonzero(s.owned)[0]
                for key, w in (("pos", 3), ("orient", 4), ("lin_mom", 3), ("ang_mom", 3)):
                    a = np.asarray(sp[key]).reshape(-1, w)[own]
                    ref = np.asarray(sw[key]).reshape(-1, w)[s.voxels[own]]
                    assert np.array_equal(a, ref), "%s differs on rank %d after chunk %d" % (key, s.rank, chunk)
        sums = np.sum([p.batch.com_sums(0) for p in parts], axis=0)
        com = sums[:3] / sums[3]
        rw = whole.results()[0]
        np.testing.assert_allclose(com, list(rw.current_com), rtol=1e-12)
        for p in parts:
            p.batch.close()
        whole.close()
    finally:
        lib.vx3_builder_destroy(b)


@pytest.mark.gpu
def test_multiprocess_ipc_halo_exchange_two_gpus():
    """One process per GPU, halo exchange through CUDA IPC peer memory (the production path); needs two devices."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "scripts", "check_decomp_mp.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0 and "DECOMP_MP OK" in r.stdout, r.stdout[-3000:]
