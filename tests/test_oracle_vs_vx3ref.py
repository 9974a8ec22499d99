"""Pins the CPU oracle's VX3 mode (oracle/vx3_oracle.cpp, the checker of every GPU parity test) on the reference's OWN VX3
code: src/VX3/*.cu + src/Utils/*.h compiled unmodified for the host (oracle/Makefile target ref_vx3, shim
oracle/ref_vx3/vxhost.h) and driven through VX3_VoxelyzeKernel::doTimeStep / CUDA_Simulation.

Every VX3-only behaviour the reference CPU library (src/old) lacks is covered: per-voxel phase actuation, the all-pairs
collision sweep, attach (link ctor, blended material, isNewLink ramp), detach, surface regeneration, static-friction
angMom = 0, force-field / attach-condition programs, cilia, signals, SecondaryExperiment, CoM / angle / closeness /
fitness, bilinear and piece-wise materials, non-zero Poisson's ratio.  The bar is BIT-EXACT on the whole state after
every chunk (same compiler, same libm, -ffp-contract=off on both sides).

Where oracle/_ref/libvxref_vx3.so is absent (GPU box without a prebuilt copy) the same scenarios are checked against the
committed fixtures tests/golden/vx3_*.json, written by tests/golden/make_golden_vx3.py from that library.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import util
from scenarios import SCENARIOS, scenario
from util import KIN, LINKF, LINKS, OracleSim, Vx3RefSim, assert_bit_equal
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200.model import expr_to_tokens

need_vx3ref = pytest.mark.skipif(not util.have_ref_vx3(), reason="oracle/_ref/libvxref_vx3.so not built (no reference tree on this box)")

STATE_KEYS = KIN + LINKF + LINKS + ["vox_flags", "temp", "vox_links", "contact_force", "link_vneg", "link_vpos", "link_axis", "link_mat",
                                    "link_strain", "link_max_strain", "link_strain_offset", "link_stress", "link_flags",
                                    "link_rest_length", "signal"]
RESULT_KEYS = ["num_voxel", "num_measured_voxel", "num_close_pairs", "steps", "num_links", "collision_count", "current_time",
               "fitness_score", "total_distance_of_all_voxels", "recent_angle", "target_closeness"]


def result_dict(r):
    out = {k: getattr(r, k) for k in RESULT_KEYS}
    out["initial_com"] = list(r.initial_com)
    out["current_com"] = list(r.current_com)
    return out


def same_number(a, b):
    return a == b or (isinstance(a, float) and a != a and b != b)


def build_scenario(name):
    sc = scenario(name)
    spec = sc["spec"]()
    lib = util.load_engine()
    b, d = spec.build(lib)
    if sc["link_capacity"]:
        d.contents.link_capacity = sc["link_capacity"]
    return sc, spec, lib, b, d


@need_vx3ref
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_bit_equals_reference_vx3(name):
    sc, spec, lib, b, d = build_scenario(name)
    try:
        ref = Vx3RefSim(spec, d)
        orc = OracleSim(d)
        assert ref.recommended_dt() == orc.recommended_dt()
        dt = float(np.float32(0.9 * orc.recommended_dt())) if sc["dt"] == "fixed" else -1.0
        done = 0
        while done < sc["steps"]:
            k = min(sc["chunk"], sc["steps"] - done)
            assert ref.step(k, dt) == k
            assert orc.step(k, dt) == k
            done += k
            sr, so = ref.state(), orc.state()
            assert_bit_equal(so, sr, STATE_KEYS, "%s after %d steps" % (name, done))
            rr, ro = result_dict(ref.result()), result_dict(orc.result())
            for key in rr:
                va, vb = (rr[key], ro[key]) if isinstance(rr[key], list) else ([rr[key]], [ro[key]])
                assert all(same_number(x, y) for x, y in zip(va, vb)), "%s after %d steps: result.%s %r != %r" % (name, done, key, ro[key], rr[key])
            assert ref.surface() == orc_surface(orc), "%s: surface voxel list" % name
    finally:
        lib.vx3_builder_destroy(b)


def orc_surface(orc):
    n = orc.counts()["n_surface"]
    out = (C.c_int * max(n, 1))()
    got = orc.lib.vx3o_surface(orc.h, out, n)
    return list(out[:got])


@need_vx3ref
def test_scenarios_reach_what_they_claim():
    """The scenarios really drive the branches they are meant to pin (checked on the reference's own state)."""
    def final(name):
        sc, spec, lib, b, d = build_scenario(name)
        try:
            ref = Vx3RefSim(spec, d)
            dt = float(np.float32(0.9 * ref.recommended_dt())) if sc["dt"] == "fixed" else -1.0
            hist = []
            for _ in range(sc["steps"] // sc["chunk"]):
                ref.step(sc["chunk"], dt)
                hist.append(ref.state())
            return ref, hist
        finally:
            lib.vx3_builder_destroy(b)
    ref, h = final("pile_sticky")
    assert h[-1]["link_vneg"].shape[0] > h[0]["link_vneg"].shape[0] or ref.counts()["n_links"] > 16, "attach must create links"
    ref, h = final("detach")
    assert any((s["link_flags"] & abi.LINKSTATE_DETACHED).any() for s in h), "detach must fire"
    ref, h = final("cantilever")
    assert ((h[-1]["link_flags"] & abi.LINKSTATE_SMALL_ANGLE) == 0).any(), "large-angle regime"
    for name in ("bilinear", "datamat", "poisson_bilinear", "poisson_data"):
        ref, h = final(name)
        assert any((s["link_strain_offset"] > 0).any() for s in h), "%s: plastic offset (yield passed)" % name
        assert any(((s["link_strain"] < s["link_max_strain"]) & (s["link_strain_offset"] > 0)).any() for s in h), "%s: unloading branch" % name
    for name in ("bilinear_fail", "datamat_fail", "linfail"):
        ref, h = final(name)
        zero = (np.abs(h[-1]["link_force_neg"]).sum(axis=1) == 0) & (h[-1]["link_max_strain"] > 0)
        assert zero.any(), "%s: a failed link carries no force" % name
    ref, h = final("sig_body")
    assert any(s["signal"][:, 0].any() for s in h)
    ref, h = final("secondary")
    assert (h[-1]["link_flags"] & abi.LINKSTATE_REMOVED).any()
    ref, h = final("closeness")
    r = ref.result()
    assert r.target_closeness > 0 and r.num_close_pairs > 0


# ---------------------------------------------------------------- math tree: all 24 operators
_TERMS = [
    ("MUL", ("SIN", ("VAR", "x")), ("COS", ("VAR", "y"))),
    ("SUB", ("TAN", ("MUL", ("VAR", "z"), ("CONST", 0.3))), ("ATAN", ("VAR", "hit"))),
    ("DIV", ("LOG", ("ADD", ("ABS", ("VAR", "t")), ("E",))), ("SQRT", ("ADD", ("PI",), ("ABS", ("VAR", "angle"))))),
    ("POW", ("ADD", ("ABS", ("VAR", "targetCloseness")), ("CONST", 1.5)), ("CONST", 0.7)),
    ("INT", ("MUL", ("VAR", "numClosePairs"), ("CONST", 0.37))),
    ("NORMALCDF", ("SUB", ("VAR", "x"), ("VAR", "y"))),
    ("AND", ("GREATERTHAN", ("VAR", "x"), ("CONST", 0.1)), ("LESSTHAN", ("VAR", "y"), ("CONST", 0.9))),
    ("OR", ("NOT", ("VAR", "hit")), ("GREATERTHAN", ("VAR", "z"), ("VAR", "num_voxel"))),
    ("MUL", ("VAR", "num_voxel"), ("CONST", 1e-3)),
]
ALL_OPS_EXPR = _TERMS[0]
for _t in _TERMS[1:]:
    ALL_OPS_EXPR = ("ADD", ALL_OPS_EXPR, _t)


def tokens_array(expr):
    toks = expr_to_tokens(expr)
    arr = (abi.Token * len(toks))()
    for i, (op, val) in enumerate(toks):
        arr[i].op, arr[i].value = op, val
    return arr, len(toks)


def test_all_ops_expression_uses_every_operator():
    arr, n = tokens_array(ALL_OPS_EXPR)
    assert {arr[i].op for i in range(n)} == set(range(24))


@need_vx3ref
def test_math_tree_all_24_ops_bit_equal_reference_eval():
    ref, orc = util.load_ref_vx3(), util.load_oracle()
    arr, n = tokens_array(ALL_OPS_EXPR)
    rng = np.random.RandomState(7)
    for trial in range(400):
        v = rng.uniform(-2, 2, 9)
        v[3] = float(rng.randint(0, 3))
        v[7] = float(rng.randint(0, 9))
        v[8] = float(rng.randint(1, 50))
        vv = (C.c_double * 9)(*v)
        a, b = ref.vx3ref_eval(arr, n, vv), orc.vx3o_eval(arr, n, vv)
        assert same_number(a, b), (trial, a, b)
    # every operator on its own, edge values included
    unary = ["SIN", "COS", "TAN", "ATAN", "LOG", "INT", "ABS", "NOT", "SQRT", "NORMALCDF"]
    binary = ["ADD", "SUB", "MUL", "DIV", "POW", "GREATERTHAN", "LESSTHAN", "AND", "OR"]
    edge = [0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 1e-300, 1e300, float("inf"), -1.0]
    for op in unary:
        for x in edge:
            arr, n = tokens_array((op, ("CONST", x)))
            vv = (C.c_double * 9)(*([0.0] * 9))
            assert same_number(ref.vx3ref_eval(arr, n, vv), orc.vx3o_eval(arr, n, vv)), (op, x)
    for op in binary:
        for x in edge:
            for y in edge:
                arr, n = tokens_array((op, ("CONST", x), ("CONST", y)))
                vv = (C.c_double * 9)(*([0.0] * 9))
                assert same_number(ref.vx3ref_eval(arr, n, vv), orc.vx3o_eval(arr, n, vv)), (op, x, y)


# ---------------------------------------------------------------- the CUDA_Simulation loop itself (stop condition, fitness)
@need_vx3ref
@pytest.mark.parametrize("name", ["runner", "closeness_run"])
def test_cuda_simulation_loop_bit_equals_oracle_run(name):
    from scenarios import closeness_spec, runner_spec
    spec = runner_spec() if name == "runner" else closeness_spec()
    if name == "closeness_run":
        spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.03)))
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        ref = Vx3RefSim(spec, d)
        out = ref.run_simulation()
        orc = OracleSim(d)
        orc.run()
        rr, ro = result_dict(ref.result(refresh=False)), result_dict(orc.result(refresh=False))
        for key in rr:
            va, vb = (rr[key], ro[key]) if isinstance(rr[key], list) else ([rr[key]], [ro[key]])
            assert all(same_number(x, y) for x, y in zip(va, vb)), "result.%s %r != %r" % (key, ro[key], rr[key])
        assert_bit_equal(orc.state(), ref.state(), STATE_KEYS, name)
        assert b"real_stepsize:" in out
    finally:
        lib.vx3_builder_destroy(b)


# ---------------------------------------------------------------- committed fixtures (travel to the GPU box)
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_matches_vx3_golden_fixture(name):
    p = os.path.join(util.GOLDEN, "vx3_" + name + ".json")
    assert os.path.exists(p), "missing fixture %s (run tests/golden/make_golden_vx3.py where /root/reference exists)" % p
    g = json.load(open(p))
    sc, spec, lib, b, d = build_scenario(name)
    try:
        orc = OracleSim(d)
        dt = float.fromhex(g["dt"])
        assert orc.step(g["steps"], dt) == g["steps"]
        so = orc.state()
        for k, hexes in g["state"].items():
            got = np.asarray(so[k], np.float64).ravel()
            want = np.array([float.fromhex(x) for x in hexes])
            assert got.shape == want.shape, (name, k)
            np.testing.assert_array_equal(got, want, err_msg="%s.%s" % (name, k))
        for k, ints in g["ints"].items():
            np.testing.assert_array_equal(np.asarray(so[k]).ravel(), np.array(ints), err_msg="%s.%s" % (name, k))
    finally:
        lib.vx3_builder_destroy(b)
