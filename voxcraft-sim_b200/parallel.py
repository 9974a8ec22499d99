"""Multi-GPU host logic: the path shards by independent simulation, exactly like the reference's worker.

Reference: vx3_node_worker assigns file i to device i % nDevices (src/Executables/vx3_node_worker.cu:88-93); each
device runs its sub-batch without any communication; the results are concatenated on the host and sorted by
fitness, NaN last (src/VX3/VX3_SimulationManager.cu:150-154,472; VX3_SimulationResult.h:26-33).

Here: one process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  The only collective is the
end-of-batch gather of one small record per simulation.
"""
import math

import torch
import torch.distributed as dist

RESULT_FIELDS = ("index", "status", "steps", "num_voxel", "current_time", "fitness", "com_x", "com_y", "com_z",
                 "com0_x", "com0_y", "com0_z", "total_distance")


def shard_indices(n_items, world, rank):
    """Indices of the simulations rank `rank` runs: i % world == rank (the reference's round-robin rule)."""
    return list(range(rank, n_items, world))


def pack_results(indices, results):
    """vx3_result records of this rank's simulations -> float64 tensor [n, len(RESULT_FIELDS)]."""
    t = torch.zeros((len(indices), len(RESULT_FIELDS)), dtype=torch.float64)
    for row, (i, r) in enumerate(zip(indices, results)):
        t[row] = torch.tensor([i, r.status, r.steps, r.num_voxel, r.current_time, r.fitness_score,
                               r.current_com[0], r.current_com[1], r.current_com[2],
                               r.initial_com[0], r.initial_com[1], r.initial_com[2], r.total_distance_of_all_voxels],
                              dtype=torch.float64)
    return t


def gather_results(local, n_total, device=None):
    """All ranks contribute their [n_local, F] block; every rank gets the [n_total, F] table in simulation order.
    Uneven shards are padded to the largest shard (all_gather needs equal shapes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.zeros((n_total, local.shape[1]), dtype=torch.float64)
        out[local[:, 0].long()] = local
        return out
    world = dist.get_world_size()
    per = int(math.ceil(n_total / world))
    dev = device if device is not None else local.device
    pad = torch.full((per, local.shape[1]), -1.0, dtype=torch.float64, device=dev)
    pad[: local.shape[0]] = local.to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.zeros((n_total, local.shape[1]), dtype=torch.float64)
    for p in parts:
        p = p.cpu()
        p = p[p[:, 0] >= 0]
        out[p[:, 0].long()] = p
    return out


def sort_by_fitness(table):
    """Row order of the report: fitness descending, NaN (diverged) last, stable (sortResults)."""
    fit = table[:, RESULT_FIELDS.index("fitness")]
    rows = list(range(table.shape[0]))
    ok = [i for i in rows if not math.isnan(float(fit[i]))]
    bad = [i for i in rows if math.isnan(float(fit[i]))]
    ok.sort(key=lambda i: -float(fit[i]))
    return table[ok + bad]
