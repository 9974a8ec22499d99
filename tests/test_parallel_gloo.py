"""World-size-2 gloo test (CPU) of the multi-GPU host logic: round-robin sharding of independent simulations and
the end-of-batch result gather + fitness sort — the only cross-device step of the path."""
import math
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as graft

graft.load_package()
from voxcraft_sim_b200 import abi, parallel  # noqa: E402

N_SIMS = 11


def fake_result(i):
    r = abi.Result()
    r.status = 2 if i % 5 == 3 else 1
    r.steps = 100 + i
    r.num_voxel = 10 * i + 1
    r.current_time = 0.5 + i
    r.fitness_score = float("nan") if i % 5 == 3 else math.sin(i * 1.7)
    for k in range(3):
        r.current_com[k] = i + 0.1 * k
        r.initial_com[k] = -i
    r.total_distance_of_all_voxels = 3.0 * i
    return r


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_indices(N_SIMS, world, rank)
    local = parallel.pack_results(mine, [fake_result(i) for i in mine])
    table = parallel.gather_results(local, N_SIMS)
    q.put((rank, mine, table.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_rule_is_reference_round_robin():
    assert parallel.shard_indices(7, 3, 0) == [0, 3, 6]
    assert parallel.shard_indices(7, 3, 2) == [2, 5]
    allidx = sorted(sum((parallel.shard_indices(4096, 8, r) for r in range(8)), []))
    assert allidx == list(range(4096))


def test_gather_and_sort_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = parallel.pack_results(list(range(N_SIMS)), [fake_result(i) for i in range(N_SIMS)])
    for rank, mine, table in got:
        assert mine == list(range(rank, N_SIMS, 2))
        t = torch.tensor(table, dtype=torch.float64)
        same = (t == serial) | (torch.isnan(t) & torch.isnan(serial))
        assert bool(same.all()), "rank %d gathered table differs from the serial one" % rank
    ranked = parallel.sort_by_fitness(serial)
    fit = ranked[:, parallel.RESULT_FIELDS.index("fitness")]
    n_ok = int((~torch.isnan(fit)).sum())
    assert all(float(fit[i]) >= float(fit[i + 1]) for i in range(n_ok - 1))
    assert bool(torch.isnan(fit[n_ok:]).all()) and n_ok == N_SIMS - len([i for i in range(N_SIMS) if i % 5 == 3])
