#!/usr/bin/env python
"""Multi-process check of the slab decomposition (run under torchrun, one rank per GPU): a small body stepped as WORLD slabs
with the CUDA-IPC halo exchange must agree bit for bit with the same body stepped whole on rank 0's GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_decomp_mp.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

graft.load_package()
import util  # noqa: E402
from util import EngineBatch, cube_spec  # noqa: E402
from voxcraft_sim_b200 import parallel  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = cube_spec((4 * world + 3, 5, 4), seed=33, actuated=True, holes=0.1, name="decomp_mp")
    lib = util.load_engine()
    b, d = spec.build(lib)
    dt = float(np.float32(0.9 * lib.vx3_model_recommended_dt(d)))
    slab = parallel.partition_slabs(d, world, rank)
    body = parallel.DecomposedBody(slab, dt, device=local)
    body.connect()
    steps = 600
    for chunk in (1, 2, 37, 160, 400):  # several stepping calls: every call ends with a collect of its own, the next one starts in the link pass
        body.step(chunk)
    st = body.batch.state(0)
    own = np.nonzero(slab.owned)[0]
    mine = (slab.voxels[own].tolist(), np.asarray(st["pos"]).reshape(-1, 3)[own].tolist(), np.asarray(st["orient"]).reshape(-1, 4)[own].tolist())
    table = [None] * world
    dist.all_gather_object(table, mine)
    com, _ = body.center_of_mass()
    ok = True
    if rank == 0:
        whole = EngineBatch([d], device=local)
        whole.set_profiling(False, use_persistent=False)
        whole.step(steps, dt)
        sw = whole.state(0)
        pos, ori = np.asarray(sw["pos"]).reshape(-1, 3), np.asarray(sw["orient"]).reshape(-1, 4)
        seen = 0
        for ids, p, q in table:
            ok &= np.array_equal(np.asarray(p), pos[ids]) and np.array_equal(np.asarray(q), ori[ids])
            seen += len(ids)
        ok &= seen == d.contents.n_voxels
        rc = whole.results()[0]
        ok &= np.allclose(com, list(rc.current_com), rtol=1e-12, atol=0)
        print("DECOMP_MP", "OK" if ok else "MISMATCH", "world", world, "voxels", seen, "steps", steps, "com", com)
    dist.barrier()
    body.batch.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
