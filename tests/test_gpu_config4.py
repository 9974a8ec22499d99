"""GPU: BASELINE config 4 AT SIZE — the 8x8x8 pile of sticky 4^3 bodies (32,768 voxels, 28,672 surface voxels, collisions +
attach + detach) — against the oracle's all-pairs sweep (4e8 pair tests per step, the reference's own O(S^2) order).

The bodies start one lattice cell apart (two vertically), so the scenario gives them initial momenta that close the gaps
within ten steps: hundreds of bodies hit each other in the same step, several hundred links are created per step and the
first impact step produces more simultaneous attach candidates than the 2,048 the round-1 resolve kernel could hold.
Checked after every chunk: link topology (ends, axis, material of every link incl. the attach-created ones, in creation
order), voxel link slots, voxel / link flags incl. the isNewLink countdown — BIT-EXACT; collision / attach / detach counts;
kinematic state and link forces within the stated gate (tests/test_gpu_scenarios.py).
"""
import ctypes as C

import numpy as np
import pytest

import util
from util import KIN, LINKF, LINKS, EngineBatch, OracleSim, gate_within_envelope
from voxcraft_sim_b200 import workloads as W

pytestmark = pytest.mark.gpu

INT_KEYS = ["vox_flags", "vox_links", "link_vneg", "link_vpos", "link_axis", "link_flags"]


def impact_pile(grid, body=4, speed=10.0, name="c4_impact"):
    spec = W.c4_spec(grid=grid, body=body, name=name)
    lib = util.load_engine()
    b, d = spec.build(lib)
    nv = d.contents.n_voxels
    ix = np.ctypeslib.as_array(d.contents.ix, shape=(nv,)).astype(int)
    iy = np.ctypeslib.as_array(d.contents.iy, shape=(nv,)).astype(int)
    iz = np.ctypeslib.as_array(d.contents.iz, shape=(nv,)).astype(int)
    mass = d.contents.voxel_mats[0].mass
    bi, bj, bk = ix // (body + 1), iy // (body + 1), iz // (body + 2)
    mom = np.zeros((nv, 3))
    mom[:, 0] = np.where(bi % 2 == 0, speed, -speed) * mass   # neighbours in x approach each other
    mom[:, 1] = np.where(bj % 2 == 0, speed, -speed) * mass
    mom[:, 2] = np.where(bk % 2 == 1, -1.5 * speed, 0.0) * mass  # every second layer drops onto the one below
    d.contents.lin_mom = mom.ctypes.data_as(C.POINTER(C.c_double))
    d.contents.link_capacity = d.contents.n_links + 6 * nv + 1024
    return lib, b, d, mom


def run_impact(grid, steps, chunk, min_cand_peak):
    lib, b, d, keep = impact_pile(grid)
    try:
        eng, orc = EngineBatch([d]), OracleSim(d)
        done = 0
        link_cap = d.contents.link_capacity
        while done < steps:
            eng.step(chunk)
            assert orc.step(chunk, -1.0) == chunk
            done += chunk
            se, so = eng.state(0, link_cap=link_cap), orc.state()
            what = "pile %s after %d steps" % (grid, done)
            assert se["link_vneg"].shape == so["link_vneg"].shape, what + ": link count %d vs %d" % (se["link_vneg"].shape[0], so["link_vneg"].shape[0])
            for k in INT_KEYS:
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s: %s" % (what, k))
            gate_within_envelope(se, so, None, ["pos", "orient"], what, rel_floor=1e-9)
            gate_within_envelope(se, so, None, ["lin_mom", "ang_mom"] + LINKF + LINKS, what, rel_floor=1e-7)
            re, ro, oc, ec = eng.results()[0], orc.result(), orc.counts(), eng.counters(0)
            assert (re.num_links, re.collision_count, re.steps) == (ro.num_links, ro.collision_count, ro.steps), what
            assert (ec["attach"], ec["detach"]) == (oc["attach"], oc["detach"]), what
        assert oc["attach"] > 100, "the scenario must create links"
        assert ec["cand_peak"] >= min_cand_peak, "candidates in one step: %d" % ec["cand_peak"]
        print(grid, "attach", oc["attach"], "detach", oc["detach"], "peak candidates in one step", ec["cand_peak"])
        # the contact phase's two-sided depth-5 neighbour search against the reference's path walk, on the glued-together pile
        mism, positives = eng.check_neighbor_search(0, n_pairs=40000, seed=7)
        assert mism == 0 and 100 < positives < 40000, (mism, positives)
        eng.close()
    finally:
        lib.vx3_builder_destroy(b)


def test_pile_of_64_bodies_through_impact_and_rebound():
    """4x4x4 bodies (4,096 voxels) for 150 steps: impact, attach, rebound, links failing in tension and detaching."""
    run_impact((4, 4, 4), steps=150, chunk=25, min_cand_peak=100)


def test_config4_full_size_32768_voxels():
    """The full 512-body pile for 24 steps (contact starts at step 10): > 2,048 attach candidates in one step."""
    run_impact((8, 8, 8), steps=24, chunk=6, min_cand_peak=2049)
