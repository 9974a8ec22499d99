"""GPU parity over every named scenario of tests/scenarios.py — the same models on which the CPU oracle is pinned bit for
bit to the reference's own VX3 code (tests/test_oracle_vs_vx3ref.py, fixtures tests/golden/vx3_*.json) — through the C ABI.

Gates (stated once, used everywhere):
  * integer state — link topology (ends, axis, material), voxel link slots, voxel / link flags incl. the isNewLink
    countdown, collision and event counts, signal state — BIT-EXACT;
  * floating-point state — element-wise |gpu - oracle| <= max(1e-9 * max|oracle|, 8 x libm envelope) — for link forces / moments also
    stiffness x 1e-9 x max|pos| (util.link_force_floors: what the position bar itself implies for a stiff link) —: the product is built
    with -fmad=false, so the only arithmetic that may differ from the oracle's is libdevice vs glibc sin / cos / acos
    (each within 1 ulp, not identically rounded); the envelope is the oracle's own spread when exactly those results are
    jittered by +-1 ulp (util.libm_envelope).  Kinematic state passes the flat 1e-9 bar on the small scenarios; the envelope
    term matters for momenta / link forces (differences of large terms) and for chaotic scenarios (hundreds of attach /
    detach events).
"""
import numpy as np
import pytest

import util
from scenarios import SCENARIOS, scenario
from util import KIN, LINKF, LINKS, EngineBatch, OracleSim, gate_within_envelope, libm_envelope
from voxcraft_sim_b200 import abi

pytestmark = pytest.mark.gpu

INT_KEYS = ["vox_flags", "vox_links", "link_vneg", "link_vpos", "link_axis", "link_flags"]
FLOAT_KEYS = KIN + LINKF + LINKS + ["link_rest_length"]
F32_EXACTISH = ["link_strain", "link_max_strain", "link_strain_offset", "link_stress", "temp"]


def run_scenario(name, persistent=True, fused=False):
    sc = scenario(name)
    spec = sc["spec"]()
    lib = util.load_engine()
    b, d = spec.build(lib)
    if sc["link_capacity"]:
        d.contents.link_capacity = sc["link_capacity"]
    try:
        orc0 = OracleSim(d)
        dt = float(np.float32(0.9 * orc0.recommended_dt())) if sc["dt"] == "fixed" else -1.0
        checkpoints = list(range(sc["chunk"], sc["steps"] + 1, sc["chunk"]))
        exact, envs = libm_envelope(d, sc["steps"], dt, checkpoints, keys=FLOAT_KEYS + F32_EXACTISH)
        eng = EngineBatch([d])
        eng.set_profiling(False, use_persistent=persistent)
        done, worst, states = 0, {}, []
        link_cap = sc["link_capacity"] or None
        orc = OracleSim(d)
        for c, so, env in zip(checkpoints, exact, envs):
            eng.step(c - done, dt) if dt > 0 else eng.step(c - done)
            assert orc.step(c - done, dt) == c - done
            done = c
            se = eng.state(0, link_cap=link_cap)
            what = "%s after %d steps" % (name, c)
            for k in INT_KEYS:
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s: %s" % (what, k))
            np.testing.assert_array_equal(se["signal"], so["signal"], err_msg=what + ": signal state")
            w = gate_within_envelope(se, so, env, FLOAT_KEYS, what, abs_floor=util.link_force_floors(d, so))
            # strain / stress / temperature are STORED as float: a 1e-16 relative difference in the double they are rounded from
            # can flip the rounding, so their floor is one float ulp at the array's scale
            w.update(gate_within_envelope(se, so, env, F32_EXACTISH, what, rel_floor=1.2e-7))
            for k, v in w.items():
                worst[k] = max(worst.get(k, 0.0), v)
            re, ro = eng.results()[0], orc.result()
            assert (re.steps, re.num_links, re.collision_count, re.num_close_pairs, re.num_measured_voxel) == \
                   (ro.steps, ro.num_links, ro.collision_count, ro.num_close_pairs, ro.num_measured_voxel), what
            np.testing.assert_allclose([re.current_time, re.target_closeness], [ro.current_time, ro.target_closeness], rtol=1e-9, atol=1e-12, err_msg=what)
            # recentAngle = acos of the normalised dot product of two successive centre-of-mass displacements (:318-333): the
            # displacements are differences of positions that agree to 1e-9 * 0.03 m, and between two samples the centre of mass of a
            # body that mostly stands still moves 1e-5 m — the ANGLE is then only good to ~1e-6 however exact the positions are
            np.testing.assert_allclose(re.recent_angle, ro.recent_angle, rtol=1e-6, atol=1e-9, err_msg=what)
            np.testing.assert_allclose(list(re.current_com) + list(re.initial_com), list(ro.current_com) + list(ro.initial_com), rtol=1e-9, atol=1e-15,
                                       err_msg=what)
            # the fitness variables x / y / z are centre-of-mass DISPLACEMENTS (:530-534): differences of two positions that each meet
            # the 1e-9 bar, so a fitness made of them is uncertain by 1e-9 * max|CoM| in absolute terms
            com_scale = max(abs(v) for v in list(ro.current_com) + list(ro.initial_com))
            np.testing.assert_allclose(re.fitness_score, ro.fitness_score, rtol=1e-9, atol=max(1e-15, 1e-9 * com_scale), err_msg=what)
            states.append(se)
        eng.close()
        print(name, "worst err/tol:", {k: "%.2g" % v for k, v in sorted(worst.items()) if v > 0.05})
        return states
    finally:
        lib.vx3_builder_destroy(b)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario_matches_oracle(name):
    run_scenario(name)


FIXED_TOPOLOGY = ["act333", "ragged", "cantilever", "forcefield", "bilinear", "bilinear_fail", "datamat", "datamat_fail", "linfail",
                  "poisson_lin", "poisson_bilinear", "poisson_data"]
BITS = KIN + LINKF + LINKS + ["link_flags", "vox_flags", "temp", "link_rest_length", "link_strain", "link_max_strain", "link_strain_offset", "link_stress"]


@pytest.mark.parametrize("name", FIXED_TOPOLOGY)
def test_three_step_paths_are_bit_identical(name, monkeypatch):
    """Persistent on-chip kernel, two-pass streaming kernels and the fused block step run the same physics code: identical
    bits on every scenario with a fixed link topology, nonlinear materials and nu != 0 included."""
    a = run_scenario(name, persistent=True)
    b = run_scenario(name, persistent=False)
    monkeypatch.setenv("VX3_FUSED", "1")
    c = run_scenario(name, persistent=False)
    for sa, sb, sc_ in zip(a, b, c):
        util.assert_bit_equal(sa, sb, BITS, name + ": persistent vs streaming")
        util.assert_bit_equal(sc_, sb, [k for k in BITS if k not in LINKF], name + ": fused vs streaming")


def test_fitness_program_with_all_24_operators():
    """vx3_batch_run with a fitness formula that uses every math-tree operator (VX3_MathTree.h:50-192) and every variable,
    against the oracle (whose evaluator is pinned bit for bit on VX3_MathTree::eval)."""
    from scenarios import closeness_spec
    from test_oracle_vs_vx3ref import ALL_OPS_EXPR
    spec = closeness_spec()
    spec.set_program(abi.PROG_FITNESS, ALL_OPS_EXPR)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.03)))
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        eng, orc = EngineBatch([d]), OracleSim(d)
        eng.run()
        orc.run()
        re, ro = eng.results()[0], orc.result(refresh=False)
        assert re.status == ro.status == abi.SIM_STOPPED and re.steps == ro.steps
        assert ro.fitness_score == ro.fitness_score and abs(ro.fitness_score) > 1e-3
        np.testing.assert_allclose(re.fitness_score, ro.fitness_score, rtol=1e-9)
        assert re.num_close_pairs == ro.num_close_pairs > 0
        eng.close()
    finally:
        lib.vx3_builder_destroy(b)


def test_force_field_program_longer_than_128_tokens():
    """Per-voxel programs may be as long as the reference's 1024-token buffers (VX3_VoxelyzeKernel.cuh:113-119): a 500-token
    force field (a sum of 80 sinusoids) against the oracle."""
    from scenarios import forcefield_spec
    from voxcraft_sim_b200.model import expr_to_tokens
    spec = forcefield_spec()
    expr = ("CONST", 0.0)
    for i in range(80):
        expr = ("ADD", expr, ("MUL", ("CONST", 2e-5 * (1 + i % 7)), ("SIN", ("MUL", ("VAR", "t"), ("CONST", 40.0 + 13.0 * i)))))
    assert 128 < len(expr_to_tokens(expr)) <= 1024
    spec.set_program(abi.PROG_FORCE_Y, expr)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc0 = OracleSim(d)
        dt = float(np.float32(0.9 * orc0.recommended_dt()))
        exact, envs = libm_envelope(d, 400, dt, [400], keys=FLOAT_KEYS)
        for persistent in (True, False):
            eng = EngineBatch([d])
            eng.set_profiling(False, use_persistent=persistent)
            eng.step(400, dt)
            se = eng.state(0)
            gate_within_envelope(se, exact[0], envs[0], FLOAT_KEYS, "long force field, persistent=%s" % persistent)
            np.testing.assert_array_equal(se["vox_flags"], exact[0]["vox_flags"])
            eng.close()
    finally:
        lib.vx3_builder_destroy(b)


def test_mixed_batch_of_colliding_simulations():
    """Several collision / attach / detach simulations in ONE batch (the worker's normal job): every simulation has its own
    candidate and failed-link lists and its own resolve CTA, so each must equal its oracle — topology bit-exact — exactly as when
    it runs alone."""
    names = ["pile_sticky", "detach", "c4small", "touch", "pile", "secondary"]
    lib = util.load_engine()
    built = []
    try:
        for n in names:
            sc = scenario(n)
            spec = sc["spec"]()
            b, d = spec.build(lib)
            if sc["link_capacity"]:
                d.contents.link_capacity = sc["link_capacity"]
            built.append((b, d, sc))
        eng = EngineBatch([d for _, d, _ in built])
        orcs = [OracleSim(d) for _, d, _ in built]
        for chunk in range(4):
            eng.step(400)
            for i, (o, (_, d, sc)) in enumerate(zip(orcs, built)):
                assert o.step(400, -1.0) == 400
                se, so = eng.state(i, link_cap=sc["link_capacity"] or None), o.state()
                what = "%s in a mixed batch after %d steps" % (names[i], 400 * (chunk + 1))
                assert se["link_vneg"].shape == so["link_vneg"].shape, what
                for k in INT_KEYS:
                    np.testing.assert_array_equal(se[k], so[k], err_msg="%s: %s" % (what, k))
                gate_within_envelope(se, so, None, ["pos", "orient"], what, rel_floor=1e-8)
                re, ro = eng.results()[i], o.result()
                assert (re.num_links, re.collision_count, re.steps) == (ro.num_links, ro.collision_count, ro.steps), what
                c = eng.counters(i)
                oc = o.counts()
                assert (c["attach"], c["detach"]) == (oc["attach"], oc["detach"]), what
        assert sum(o.counts()["attach"] for o in orcs) > 0 and sum(o.counts()["detach"] for o in orcs) > 0
        eng.close()
    finally:
        for b, _, _ in built:
            lib.vx3_builder_destroy(b)
