"""Fused step (vx3_fused.cuh: spatial blocks, interior links + voxel integration in one CTA, face links on the two-pass
route) against the two-pass streaming kernels: the same arithmetic on the same inputs, so every array must agree BIT FOR
BIT — on small bodies cut into many small blocks (VX3_FUSE_BV), on a mixed batch, with externals / force fields /
signals + cilia, and at BASELINE sizes (config 3 sample, config 5 slice).  The oracle comparison of the fused path is in
test_gpu_parity.py; here one fused run is also held against the oracle.  The fused step is opt-in (VX3_FUSED=1)."""
import os

import numpy as np
import pytest

import util
from util import KIN, LINKF, LINKS, EngineBatch, cube_spec
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200 import workloads as W
from voxcraft_sim_b200.model import ModelSpec

pytestmark = pytest.mark.gpu

BITS = KIN + LINKF + LINKS + ["link_flags", "vox_flags", "temp", "link_rest_length", "link_strain", "link_max_strain", "link_stress"]


def run_both(descs, steps, monkeypatch, bv=None, sims=None, chunks=1):
    """[(state, result) per sim] for the fused path and for the two-pass path; also returns the fused plan's info."""
    if bv is not None:
        monkeypatch.setenv("VX3_FUSE_BV", str(bv))
    out, info = [], None
    for fused in (True, False):
        if fused:  # opt-in at creation: block storage order + fused kernels; otherwise the model's order + two-pass kernels
            monkeypatch.setenv("VX3_FUSED", "1")
        else:
            monkeypatch.delenv("VX3_FUSED", raising=False)
        eng = EngineBatch(descs)
        eng.set_profiling(False, use_persistent=False)
        if fused:
            info = eng.fused_info()
        else:
            assert eng.fused_info()[0] == 0
        for _ in range(chunks):
            eng.step(steps // chunks)
        res = eng.results()
        out.append([(eng.state(i), res[i]) for i in (sims if sims is not None else range(len(descs)))])
        eng.close()
    return out[0], out[1], info


def assert_same(fused, twopass, what):
    for k, ((sf, rf), (st, rt)) in enumerate(zip(fused, twopass)):
        util.assert_bit_equal(sf, st, BITS, "%s, simulation %d: fused vs two-pass" % (what, k))
        assert rf.steps == rt.steps and rf.current_time == rt.current_time and rf.status == rt.status
        np.testing.assert_allclose(list(rf.current_com), list(rt.current_com), rtol=1e-12)  # (summed in storage order)


@pytest.mark.parametrize("bv", [8, 27, 120])
def test_actuated_ragged_body_many_blocks(bv, monkeypatch):
    spec = cube_spec((7, 6, 5), seed=17, actuated=True, holes=0.15, lift=1, name="fused_ragged")
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        f, t, info = run_both([d], 600, monkeypatch, bv=bv, chunks=3)
        assert info[0] == 1 and info[2] + info[3] == d.contents.n_links
        if bv == 8:
            assert info[1] >= 20 and info[3] > 0  # many blocks, face links on the pre-pass
        assert_same(f, t, "ragged body, blocks of <= %d" % bv)
    finally:
        lib.vx3_builder_destroy(b)


def test_cantilever_large_angle_externals(monkeypatch):
    """The passive cantilever of test_gpu_parity.py (fixed root, end load: links leave the small-angle regime), in two blocks."""
    spec = ModelSpec(0.01, "fused_cantilever")
    spec.add_material(elastic_mod=2e5, density=1e3, u_static=1.0, u_dynamic=0.5)
    spec.set_env(bond_damping_z=0.5, col_damping_z=0.8, slow_damping_z=0.02, floor_enabled=0)
    spec.set_options(enable_collision=0)
    spec.set_structure(np.ones((1, 1, 10), np.uint8))
    spec.set_external(0, dof_fixed=0x3F)
    spec.set_external(9, force=(0.0, 0.0, -0.02))
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        f, t, info = run_both([d], 6000, monkeypatch, bv=8, chunks=2)
        assert info[0] == 1 and info[1] == 2 and info[3] == 1  # two blocks, the link between them on the pre-pass
        assert not (f[0][0]["link_flags"] & abi.LINKSTATE_SMALL_ANGLE).all(), "the load should drive links out of the small-angle regime"
        assert_same(f, t, "cantilever")
    finally:
        lib.vx3_builder_destroy(b)


def test_fused_against_the_oracle(monkeypatch):
    from util import OracleSim, compare_states
    monkeypatch.setenv("VX3_FUSED", "1")
    monkeypatch.setenv("VX3_FUSE_BV", "12")
    spec = cube_spec((5, 4, 3), seed=5, actuated=True, lift=2, holes=0.25, name="fused_oracle")
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        eng = EngineBatch([d])
        eng.set_profiling(False, use_persistent=False)
        assert eng.fused_info()[0] == 1 and eng.fused_info()[1] > 3
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        for _ in range(4):
            eng.step(300, dt)
            assert orc.step(300, dt) == 300
            se, so = eng.state(0), orc.state()
            compare_states(se, so, KIN, 1e-9, "fused vs oracle")
            compare_states(se, so, LINKF + LINKS, 1e-7, "fused vs oracle")
            np.testing.assert_array_equal(se["vox_flags"], so["vox_flags"])
            np.testing.assert_array_equal(se["link_flags"], so["link_flags"])
        re, ro = eng.results()[0], orc.result()
        assert re.steps == ro.steps
        np.testing.assert_allclose(list(re.current_com), list(ro.current_com), rtol=1e-10)
        eng.close()
    finally:
        lib.vx3_builder_destroy(b)


def test_force_field_cilia_in_one_batch(monkeypatch):
    import test_gpu_signals as TS
    cil = TS.body_spec((5, 4, 3), cilia=True, name="fb1")
    cil.set_options(enable_collision=0, enable_signals=0, enable_cilia=1)  # (signals keep the model's order: two-pass path)
    specs = [cube_spec((4, 4, 3), seed=3, actuated=True, name="fb0"), cil,
             cube_spec((3, 2, 2), seed=41, actuated=False, lift=1, name="fb2"), cube_spec((6, 5, 4), seed=9, actuated=True, holes=0.3, name="fb3")]
    specs[2].set_program(abi.PROG_FORCE_X, ("MUL", ("CONST", 1e-3), ("SIN", ("MUL", ("VAR", "t"), ("CONST", 300.0)))))
    specs[2].set_program(abi.PROG_FORCE_Z, ("MUL", ("CONST", -2e-2), ("VAR", "z")))
    specs[2].set_program(abi.PROG_ATTACH_0, ("SUB", ("VAR", "x"), ("CONST", 0.011)))
    lib = util.load_engine()
    built = [s.build(lib) for s in specs]
    try:
        f, t, info = run_both([d for _, d in built], 500, monkeypatch, bv=16, chunks=2)
        assert info[0] == 1 and info[1] > len(specs)
        assert_same(f, t, "mixed batch")
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)


def test_collisions_and_signals_keep_the_two_pass_path(monkeypatch):
    import test_gpu_signals as TS
    monkeypatch.setenv("VX3_FUSED", "1")
    lib = util.load_engine()
    for spec in (cube_spec((3, 3, 3), seed=2, actuated=True, collisions=1, name="fused_off"), TS.body_spec((4, 3, 3), name="fused_off_sig")):
        b, d = spec.build(lib)
        try:
            eng = EngineBatch([d])
            assert eng.fused_info()[0] == 0
            eng.close()
        finally:
            lib.vx3_builder_destroy(b)


def test_config3_sample_and_config5_slice(monkeypatch):
    lib = util.load_engine()
    built = [W.c3_spec(k).build(lib) for k in range(40)]
    try:
        f, t, info = run_both([d for _, d in built], 300, monkeypatch, sims=[0, 7, 21, 39])
        assert info[0] == 1
        print("config 3 sample: blocks %d, interior links %d, face links %d" % info[1:])
        assert_same(f, t, "config 3 sample")
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)
    b, d = W.c5_spec((24, 200, 100)).build(lib)
    try:
        f, t, info = run_both([d], 40, monkeypatch)
        print("config 5 slice: blocks %d, interior links %d, face links %d" % info[1:])
        assert info[0] == 1 and info[3] < 0.35 * d.contents.n_links
        assert_same(f, t, "config 5 slice")
        assert np.isfinite(f[0][0]["pos"]).all()
    finally:
        lib.vx3_builder_destroy(b)
