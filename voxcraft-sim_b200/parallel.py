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


# ------------------------------------------------------------------------------------------------------------------
# Slab decomposition of ONE body over the GPUs of a box (BASELINE config 5; new capability, not in the reference).
# Host logic only: which voxels / links each rank holds and which pose records cross each face.  The exchange itself
# runs inside the engine's step stream (csrc/engine/vx3_halo.cuh).
# ------------------------------------------------------------------------------------------------------------------
import ctypes as _C

import numpy as _np

from . import abi as _abi

_VOX_FIELDS = [("ix", _np.int16, 1), ("iy", _np.int16, 1), ("iz", _np.int16, 1), ("vox_mat", _np.int32, 1), ("pos", _np.float64, 3),
               ("orient", _np.float64, 4), ("lin_mom", _np.float64, 3), ("ang_mom", _np.float64, 3), ("vox_flags", _np.int32, 1),
               ("temp", _np.float32, 1), ("phase_offset", _np.float64, 1), ("vox_ext", _np.int32, 1), ("base_cilia", _np.float64, 3),
               ("shift_cilia", _np.float64, 3)]
_LINK_FIELDS = [("link_axis", _np.int32, 1), ("link_mat", _np.int32, 1), ("link_pos2", _np.float64, 3), ("link_angle1v", _np.float64, 3),
                ("link_angle2v", _np.float64, 3), ("link_strain", _np.float32, 1), ("link_max_strain", _np.float32, 1),
                ("link_strain_offset", _np.float32, 1), ("link_stress", _np.float32, 1), ("link_flags", _np.int32, 1),
                ("link_small_angle", _np.int32, 1), ("link_rest_length", _np.float64, 1), ("link_transverse_area", _np.float32, 1),
                ("link_transverse_strain_sum", _np.float32, 1), ("link_strain_ratio", _np.float32, 1)]


def _view(ptr, n, dtype, width=1):
    if not ptr or n == 0:
        return None
    a = _np.ctypeslib.as_array(ptr, shape=(n * width,))
    a = a.view(dtype) if a.dtype != dtype else a
    return a.reshape(n, width) if width > 1 else a


def slab_bounds(coord, world):
    """Cut positions along one lattice axis that give every rank about the same number of voxels: rank r owns the
    voxels with bounds[r] <= coord < bounds[r+1]."""
    lo, hi = int(coord.min()), int(coord.max()) + 1
    cum = _np.concatenate([[0], _np.cumsum(_np.bincount(coord - lo, minlength=hi - lo))])
    total = cum[-1]
    bounds = [lo]
    for r in range(1, world):
        k = int(_np.argmin(_np.abs(cum - total * r / world)))
        bounds.append(max(lo + k, bounds[-1] + 1))  # every slab at least one layer thick
    bounds.append(hi)
    if bounds[-2] >= hi:
        raise ValueError("body has fewer lattice layers along the split axis than there are ranks")
    return bounds


def slab_owner(desc, world, axis=0):
    """Owning rank of every voxel of the full model under partition_slabs' rule (host-side check helper)."""
    d = desc.contents if hasattr(desc, "contents") else desc
    coord = _view([d.ix, d.iy, d.iz][axis], d.n_voxels, _np.int16).astype(_np.int64)
    return _np.searchsorted(_np.asarray(slab_bounds(coord, world)[1:-1]), coord, side="right")


class SlabModel:
    """One rank's sub-model of a decomposed body: ``desc`` (a vx3_model_desc whose arrays this object keeps alive), the
    global ids of its voxels (``voxels``) and which of them it owns, and per neighbour the pose records to send / receive
    (indices into the sub-model, in the order both ranks agree on: ascending global voxel id)."""

    def __init__(self):
        self.desc = None
        self.keep = []
        self.voxels = None
        self.owned = None
        self.send = {0: _np.zeros(0, _np.int32), 1: _np.zeros(0, _np.int32)}
        self.recv = {0: _np.zeros(0, _np.int32), 1: _np.zeros(0, _np.int32)}
        self.rank = 0
        self.world = 1


def partition_slabs(desc, world, rank, axis=0):
    """Sub-model of rank ``rank``: the voxels of its slab, GHOST copies of the neighbour slabs' face voxels, every link
    with an owned end (a link crossing a face is held by both neighbours).  ``desc`` = pointer to the full vx3_model_desc."""
    d = desc.contents if hasattr(desc, "contents") else desc
    nv, nl = d.n_voxels, d.n_links
    coord = _view([d.ix, d.iy, d.iz][axis], nv, _np.int16).astype(_np.int64)
    bounds = slab_bounds(coord, world)
    owner = _np.searchsorted(_np.asarray(bounds[1:-1]), coord, side="right")
    vneg = _view(d.link_vneg, nl, _np.int32) if nl else _np.zeros(0, _np.int32)
    vpos = _view(d.link_vpos, nl, _np.int32) if nl else _np.zeros(0, _np.int32)
    own = owner == rank
    on, op = owner[vneg], owner[vpos]
    if nl and _np.abs(on - op).max() > 1:
        raise ValueError("a link joins two non-adjacent slabs: slabs must be at least one lattice layer thick")
    keep_l = _np.nonzero((on == rank) | (op == rank))[0]
    # links with two owned ends first, links across a face last: the engine evaluates the first range while the halo exchange of
    # the previous step is still in flight (csrc/engine/vx3_halo.cuh) — the order of a model's links carries no meaning
    face_l = on[keep_l] != op[keep_l]
    keep_l = _np.concatenate([keep_l[~face_l], keep_l[face_l]])
    used = own.copy()
    used[vneg[keep_l]] = True
    used[vpos[keep_l]] = True
    keep_v = _np.nonzero(used)[0]
    # owned voxels first, ghosts last: the engine's voxel pass runs over the owned prefix only (a ghost's record arrives with the halo
    # exchange; the order of a model's voxels carries no meaning where slabs are allowed: no collisions, no signals)
    keep_v = _np.concatenate([keep_v[own[keep_v]], keep_v[~own[keep_v]]])
    newv = -_np.ones(nv, _np.int32)
    newv[keep_v] = _np.arange(len(keep_v), dtype=_np.int32)
    newl = -_np.ones(max(nl, 1), _np.int32)
    newl[keep_l] = _np.arange(len(keep_l), dtype=_np.int32)

    m = SlabModel()
    m.rank, m.world = rank, world
    m.voxels = keep_v
    m.owned = own[keep_v]
    out = _abi.ModelDesc()
    _C.memmove(_C.byref(out), _C.byref(d), _C.sizeof(_abi.ModelDesc))  # palette, externals, options, programs: shared
    out.n_voxels, out.n_links, out.link_capacity = len(keep_v), len(keep_l), 0

    def put(field, arr, ctype):
        arr = _np.ascontiguousarray(arr)
        m.keep.append(arr)
        setattr(out, field, arr.ctypes.data_as(_C.POINTER(ctype)))

    ctypes_of = {_np.int16: _C.c_int16, _np.int32: _C.c_int32, _np.float32: _C.c_float, _np.float64: _C.c_double}
    for name, dt, w in _VOX_FIELDS:
        src = _view(getattr(d, name), nv, dt, w)
        if src is None:
            continue
        a = src[keep_v].copy()
        if name == "vox_flags":
            a[~m.owned] |= _abi.VOX_GHOST
        put(name, a, ctypes_of[dt])
    vl = _view(d.vox_links, nv, _np.int32, 6)[keep_v].copy()
    vl[vl >= 0] = newl[vl[vl >= 0]]  # links without an owned end drop out of the ghosts' slots
    put("vox_links", vl, _C.c_int32)
    put("link_vneg", newv[vneg[keep_l]], _C.c_int32)
    put("link_vpos", newv[vpos[keep_l]], _C.c_int32)
    for name, dt, w in _LINK_FIELDS:
        src = _view(getattr(d, name), nl, dt, w)
        if src is not None:
            put(name, src[keep_l].copy(), ctypes_of[dt])
    m.desc = out

    # face lists: a link whose ends have different owners makes each end a ghost on the other side
    cross = _np.nonzero(on != op)[0]
    for side, nb in ((0, rank - 1), (1, rank + 1)):
        if nb < 0 or nb >= world:
            continue
        mine_n = cross[(on[cross] == rank) & (op[cross] == nb)]   # my end is the negative one
        mine_p = cross[(op[cross] == rank) & (on[cross] == nb)]   # my end is the positive one
        send_g = _np.unique(_np.concatenate([vneg[mine_n], vpos[mine_p]]))
        recv_g = _np.unique(_np.concatenate([vpos[mine_n], vneg[mine_p]]))
        m.send[side] = newv[send_g].astype(_np.int32)
        m.recv[side] = newv[recv_g].astype(_np.int32)
    return m


class DecomposedBody:
    """One rank's part of a body stepped over several GPUs: engine batch of the slab + the halo wiring.  ``connect()`` swaps
    the CUDA IPC handles through torch.distributed (one small all_gather of Python objects); after it, ``step`` needs no
    host communication at all."""

    def __init__(self, slab, dt, device=0, fma=False):
        from .engine import Batch
        self.slab, self.dt = slab, float(dt)
        self.batch = Batch([_C.pointer(slab.desc)], fma=fma, device=device)
        for side in (0, 1):
            if len(slab.send[side]) or len(slab.recv[side]):
                self.batch.halo_setup(side, slab.send[side], slab.recv[side])

    def connect(self):
        info = {}
        for side in (0, 1):
            if len(self.slab.send[side]) or len(self.slab.recv[side]):
                info[side] = (self.batch.halo_export(side), int(len(self.slab.recv[side])))
        table = [None] * dist.get_world_size()
        dist.all_gather_object(table, info)
        r = self.slab.rank
        for side, nb in ((0, r - 1), (1, r + 1)):
            if side in info and 0 <= nb < self.slab.world:
                handle, n_recv = table[nb][1 - side]  # the neighbour's block for ITS side facing me
                self.batch.halo_connect(side, handle, n_recv)

    def step(self, k):
        self.batch.step(k, self.dt)

    def center_of_mass(self):
        """Global centre of mass: per-rank sums over owned voxels, added over the ranks (the one collective of the path)."""
        s = torch.tensor(self.batch.com_sums(0), dtype=torch.float64)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":
                s = s.cuda()
            dist.all_reduce(s)
            s = s.cpu()
        return [float(s[k] / s[3]) if s[3] != 0 else 0.0 for k in range(3)], s
