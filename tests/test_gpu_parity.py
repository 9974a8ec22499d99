"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle — behaviours that need their own driver code
(stop conditions, divergence, batches, arena reuse, grid overflow).  The scenario-by-scenario state parity, with the
tolerance gates stated and justified, is tests/test_gpu_scenarios.py.

The product library is compiled with -fmad=false, so every fp64/fp32 operation rounds exactly like the oracle's (and the
reference's) non-contracting x86-64 build; the only arithmetic that can differ is libdevice vs glibc sin/cos/acos (<= 1 ulp).
"""
import ctypes as C

import numpy as np
import pytest

import util
from scenarios import collide_spec
from util import KIN, LINKF, LINKS, EngineBatch, OracleSim, compare_states, cube_spec
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200.model import ModelSpec

pytestmark = pytest.mark.gpu

TOL_KIN = 1e-9
TOL_LINK = 1e-7


def build(spec):
    lib = util.load_engine()
    b, d = spec.build(lib)
    return lib, b, d


def check_state(se, so, what, kin=TOL_KIN, link=TOL_LINK, links=True):
    w = compare_states(se, so, KIN, kin, what)
    if links:
        w.update(compare_states(se, so, LINKF + LINKS, link, what))
    return w


def run_pair(spec, steps, dt_scale=0.9, links=True, check_every=None, persistent=True):
    """persistent=True lets the engine pick the on-chip persistent kernel where it applies (single collision-free
    body); False forces the streaming kernels; "mixed" alternates between the two from chunk to chunk."""
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        eng.set_profiling(False, use_persistent=bool(persistent))
        orc = OracleSim(d)
        dt = float(np.float32(dt_scale * orc.recommended_dt()))
        chunk = check_every or steps
        done = 0
        worst = {}
        while done < steps:
            k = min(chunk, steps - done)
            if persistent == "mixed":
                eng.set_profiling(False, use_persistent=(done // chunk) % 2 == 0)
            eng.step(k, dt)
            assert orc.step(k, dt) == k
            done += k
            se, so = eng.state(0), orc.state()
            w = check_state(se, so, "%s step %d" % (spec.name, done), links=links)
            np.testing.assert_array_equal(se["vox_flags"], so["vox_flags"])
            np.testing.assert_array_equal(se["link_flags"], so["link_flags"])
            for kk, vv in w.items():
                worst[kk] = max(worst.get(kk, 0.0), vv)
        re, ro = eng.results()[0], orc.result()
        assert re.steps == ro.steps
        assert abs(re.current_time - ro.current_time) <= 1e-15 * max(1.0, abs(ro.current_time))
        eng.close()
        print(spec.name, {k: "%.1e" % v for k, v in worst.items()})
        return worst
    finally:
        lib.vx3_builder_destroy(b)


def test_single_voxel_drop():
    """demos/basic analogue: one passive voxel falling onto the floor."""
    spec = ModelSpec(0.01, "drop")
    spec.add_material(elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.5)
    st = np.zeros((3, 1, 1), np.uint8)
    st[2, 0, 0] = 1
    spec.set_structure(st)
    run_pair(spec, 3000, links=False, check_every=500)


@pytest.mark.parametrize("persistent", [True, False, "mixed"])
@pytest.mark.parametrize("shape,steps", [((2, 1, 1), 1000), ((3, 3, 3), 1000), ((6, 6, 6), 400)])
def test_actuated_body(shape, steps, persistent):
    spec = cube_spec(shape, seed=11, actuated=True, name="act%dx%dx%d" % shape)
    run_pair(spec, steps, check_every=100, persistent=persistent)


def test_persistent_equals_streaming_bitwise():
    """The on-chip persistent kernel and the streaming kernels run the same arithmetic: identical bits after 700 steps."""
    spec = cube_spec((7, 6, 5), seed=17, actuated=True, holes=0.1, name="bitwise")
    lib, b, d = build(spec)
    try:
        out = []
        for use in (True, False):
            eng = EngineBatch([d])
            eng.set_profiling(False, use_persistent=use)
            eng.step(700)
            out.append(eng.state(0))
            eng.close()
        util.assert_bit_equal(out[0], out[1], KIN + LINKF + LINKS + ["link_flags", "vox_flags", "temp", "link_rest_length", "link_strain",
                                                                      "link_max_strain", "link_stress"], "persistent vs streaming")
    finally:
        lib.vx3_builder_destroy(b)



@pytest.mark.parametrize("persistent", [True, False])
def test_divergence_stops_at_the_reference_step(persistent):
    """A bar torn apart by a huge end load: some link's strain passes 100, doTimeStep returns false in that step
    (VX3_VoxelyzeKernel.cu:273-281) — CurStepCount counts it, currentTime does not, the fitness reads NaN.  Both GPU paths
    must stop at the oracle's step."""
    spec = ModelSpec(0.01, "tear")
    spec.add_material(elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.5)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01, floor_enabled=0, grav_enabled=0)
    spec.set_options(enable_collision=0)
    spec.set_structure(np.ones((2, 3, 12), np.uint8))
    for i in range(6):
        spec.set_external(12 * i, dof_fixed=0x3F)
        spec.set_external(12 * i + 11, force=(6e3, 0.0, 0.0))  # diverges after ~130 steps
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        eng.set_profiling(False, use_persistent=persistent)
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        eng.step(5000, dt)
        done = orc.step(5000, dt)
        assert done < 5000, "scenario must diverge"
        re, ro = eng.results()[0], orc.result()
        assert ro.status == abi.SIM_DIVERGED and re.status == abi.SIM_DIVERGED
        assert re.steps == ro.steps == done + 1
        assert re.current_time == ro.current_time
        assert np.isnan(re.fitness_score)
        # doTimeStep returned false BEFORE any voxel moved in that step: the state is the oracle's, on the persistent path too
        # (its CTAs run up to a step apart; the engine replays a diverged stretch on the streaming path, vx3_engine.cu persist_guard)
        se, so = eng.state(0), orc.state()
        compare_states(se, so, KIN, 1e-9, "state of the diverged simulation")
        np.testing.assert_array_equal(se["link_flags"], so["link_flags"])
        eng.step(10, dt)  # a finished simulation does not move
        assert eng.results()[0].steps == ro.steps
        util.assert_bit_equal(eng.state(0), se, KIN, "a finished simulation does not move")
    finally:
        lib.vx3_builder_destroy(b)


def test_batch_of_different_bodies():
    """Several simulations in one batch advance independently and equal their single runs."""
    specs = [cube_spec((3, 3, 3), seed=21, name="b0"), cube_spec((4, 2, 3), seed=22, holes=0.2, name="b1"),
             cube_spec((2, 2, 5), seed=23, lift=1, name="b2"), cube_spec((1, 1, 1), seed=24, lift=1, name="b3")]
    lib = util.load_engine()
    built = [s.build(lib) for s in specs]
    try:
        eng = EngineBatch([d for _, d in built])
        eng.step(500)
        for i, (_, d) in enumerate(built):
            orc = OracleSim(d)
            orc.step(500, -1.0)
            check_state(eng.state(i), orc.state(), "batch sim %d" % i)
            re, ro = eng.results()[i], orc.result()
            assert re.steps == ro.steps == 500
            assert re.dt == ro.dt
            np.testing.assert_allclose(list(re.current_com), list(ro.current_com), rtol=1e-12, atol=1e-15)
        # per-voxel outputs of collectResults: one simulation at a time, and the whole batch in one call (sim = -1)
        parts = [eng.positions(i) for i in range(len(built))]
        ip_all, p_all, m_all = eng.positions(None)
        np.testing.assert_array_equal(ip_all, np.concatenate([p[0] for p in parts]))
        np.testing.assert_array_equal(p_all, np.concatenate([p[1] for p in parts]))
        np.testing.assert_array_equal(m_all, np.concatenate([p[2] for p in parts]))
        np.testing.assert_array_equal(parts[1][1], eng.state(1)["pos"])
        assert set(m_all) <= {1, 2, 3}
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)


def test_batch_recreated_on_the_cached_arena_is_identical():
    """Destroying a batch returns its arena / staging buffer to the cache; a batch created on the recycled (dirty) arena
    must behave exactly like one created on fresh memory."""
    specs = [cube_spec((4, 3, 3), seed=31, name="r0"), cube_spec((3, 3, 4), seed=32, holes=0.2, name="r1")]
    lib = util.load_engine()
    built = [s.build(lib) for s in specs]
    try:
        ref = None
        for rounds in range(3):
            eng = EngineBatch([d for _, d in built])
            eng.step(300)
            st = [eng.state(i) for i in range(2)]
            eng.close()
            if ref is None:
                ref = st
                big = EngineBatch([built[0][1]] * 1 + [built[1][1]])  # another user of the cache in between
                big.step(50)
                big.close()
            else:
                for i in range(2):
                    util.assert_bit_equal(st[i], ref[i], KIN + LINKF + LINKS + ["link_flags", "vox_flags"], "recreated batch, sim %d" % i)
        lib.vx3_engine_trim()
        eng = EngineBatch([d for _, d in built])
        eng.step(300)
        util.assert_bit_equal(eng.state(0), ref[0], KIN, "after trim")
        eng.close()
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)


def test_run_stop_condition_and_fitness():
    """vx3_batch_run: stop at t > 0.05 s, fitness = sqrt(x^2+y^2) of the CoM displacement, CoM sampling cadence."""
    spec = cube_spec((3, 3, 2), seed=31, actuated=True, name="runner")
    spec.set_env(temp_period=0.01)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.05)))
    spec.set_program(abi.PROG_FITNESS, ("SQRT", ("ADD", ("MUL", ("VAR", "x"), ("VAR", "x")), ("MUL", ("VAR", "y"), ("VAR", "y")))))
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        eng.run(steps_per_launch=37)
        orc.run()
        re, ro = eng.results()[0], orc.result(refresh=False)
        assert re.status == ro.status == abi.SIM_STOPPED
        assert re.steps == ro.steps
        assert re.current_time == ro.current_time
        np.testing.assert_allclose(re.fitness_score, ro.fitness_score, rtol=1e-9)
        np.testing.assert_allclose(list(re.current_com), list(ro.current_com), rtol=1e-10)
        np.testing.assert_allclose(re.recent_angle, ro.recent_angle, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(re.total_distance_of_all_voxels, ro.total_distance_of_all_voxels, rtol=1e-9)
        assert re.num_measured_voxel == ro.num_measured_voxel
    finally:
        lib.vx3_builder_destroy(b)








def test_collision_grid_overflow_chains(monkeypatch):
    """A 4-bucket hash table: every bucket holds many cells' voxels, far more than its 8 inline slots, so the walk goes
    through the overflow chains and meets mostly foreign voxels.  Nothing may change: same contacts, same attach events."""
    monkeypatch.setenv("VX3_GRID_BUCKETS", "4")
    spec = collide_spec(True, name="overflow")
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        for i in range(8):
            eng.step(500, dt)
            orc.step(500, dt)
            se, so = eng.state(0), orc.state()
            for k in ("link_vneg", "link_vpos", "link_axis", "vox_links", "link_flags", "vox_flags"):
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s chunk %d" % (k, i))
            check_state(se, so, "overflow chunk %d" % i)
        assert orc.counts()["attach"] > 0
        assert eng.results()[0].collision_count == orc.result().collision_count > 0
    finally:
        lib.vx3_builder_destroy(b)








def test_no_device_side_cpu_fallback_symbols():
    """The product library must not contain the oracle: its only physics lives in CUDA kernels."""
    lib = util.load_engine()
    assert not hasattr(lib, "vx3o_step")
