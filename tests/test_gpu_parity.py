"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded models.

The product library is compiled with -fmad=false, so every fp64/fp32 operation rounds exactly like the oracle's
(and the reference's) non-contracting x86-64 build; the only arithmetic that can differ is libdevice vs glibc
sin/cos/acos (<= 1 ulp).  Tolerances are per-array max-norm relative errors:
  * kinematic state (pos, orient, linMom, angMom): 1e-9   — SURVEY.md §8(d)'s bar
  * link forces / moments and link-local state:     1e-7   — derived quantities; the reference's beam shear
    b1*y - b2*(a1+a2) cancels to a small net, so the same 1-ulp seed shows ~100x larger there
  * topology, flags, event counts: bit-exact.
The FMA-contracted variant (libvx3_b200_fma.so) is NOT parity-grade: the reference's formulas amplify a
contraction difference ~1e6-fold (1 - w*w in Quat3D::ToRotationVector for small angles, float casts of strain);
it is only checked loosely (positions 1e-4) — the same gap separates the reference's own nvcc and CPU builds.
"""
import ctypes as C

import numpy as np
import pytest

import util
from util import KIN, LINKF, LINKS, EngineBatch, OracleSim, compare_states, cube_spec
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200.model import ModelSpec

pytestmark = pytest.mark.gpu

TOL_KIN = 1e-9
TOL_LINK = 1e-7


def build(spec, fma=False):
    lib = util.load_engine(fma)
    b, d = spec.build(lib)
    return lib, b, d


def check_state(se, so, what, kin=TOL_KIN, link=TOL_LINK, links=True):
    w = compare_states(se, so, KIN, kin, what)
    if links:
        w.update(compare_states(se, so, LINKF + LINKS, link, what))
    return w


def run_pair(spec, steps, fma=False, dt_scale=0.9, links=True, check_every=None, persistent=True):
    """persistent=True lets the engine pick the on-chip persistent kernel where it applies (single collision-free
    body); False forces the streaming kernels; "mixed" alternates between the two from chunk to chunk."""
    lib, b, d = build(spec, fma)
    try:
        eng = EngineBatch([d], fma=fma)
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
            if fma:
                w = compare_states(se, so, ["pos"], 1e-4, "%s (fma) step %d" % (spec.name, done))
            else:
                w = check_state(se, so, "%s step %d" % (spec.name, done), links=links)
                np.testing.assert_array_equal(se["vox_flags"], so["vox_flags"])
                np.testing.assert_array_equal(se["link_flags"], so["link_flags"])
            for kk, vv in w.items():
                worst[kk] = max(worst.get(kk, 0.0), vv)
        re, ro = eng.results()[0], orc.result()
        assert re.steps == ro.steps
        assert abs(re.current_time - ro.current_time) <= 1e-15 * max(1.0, abs(ro.current_time))
        eng.close()
        print(spec.name, "fma" if fma else "product", {k: "%.1e" % v for k, v in worst.items()})
        return worst
    finally:
        lib.vx3_builder_destroy(b)


@pytest.mark.parametrize("fma", [False, True])
def test_single_voxel_drop(fma):
    """demos/basic analogue: one passive voxel falling onto the floor."""
    spec = ModelSpec(0.01, "drop")
    spec.add_material(elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.5)
    st = np.zeros((3, 1, 1), np.uint8)
    st[2, 0, 0] = 1
    spec.set_structure(st)
    run_pair(spec, 3000, fma, links=False, check_every=500)


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("persistent", [True, False, "mixed"])
@pytest.mark.parametrize("shape,steps", [((2, 1, 1), 1000), ((3, 3, 3), 1000), ((6, 6, 6), 400)])
def test_actuated_body(shape, steps, fma, persistent):
    spec = cube_spec(shape, seed=11, actuated=True, name="act%dx%dx%d" % shape)
    run_pair(spec, steps, fma, check_every=100, persistent=persistent)


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


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("persistent", [True, False])
def test_body_with_holes_and_lift(fma, persistent):
    """Ragged lattice (random holes), dropped from 2 voxels up: exercises missing links, free fall, floor contact, friction."""
    spec = cube_spec((5, 4, 3), seed=5, actuated=True, lift=2, holes=0.25, name="ragged")
    run_pair(spec, 1500, fma, check_every=250, persistent=persistent)


def test_passive_large_angle():
    """A passive cantilever with a fixed root and an end load: drives links out of the small-angle regime."""
    spec = ModelSpec(0.01, "cantilever")
    spec.add_material(elastic_mod=2e5, density=1e3, u_static=1.0, u_dynamic=0.5)
    spec.set_env(bond_damping_z=0.5, col_damping_z=0.8, slow_damping_z=0.02, floor_enabled=0)
    spec.set_structure(np.ones((1, 1, 10), np.uint8))
    spec.set_external(0, dof_fixed=0x3F)
    spec.set_external(9, force=(0.0, 0.0, -0.02))
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        eng.step(6000, dt)
        orc.step(6000, dt)
        se, so = eng.state(0), orc.state()
        small = (so["link_flags"] & abi.LINKSTATE_SMALL_ANGLE) != 0
        assert (~small).any(), "test must reach the large-angle branch"
        check_state(se, so, "cantilever")
        np.testing.assert_array_equal(se["link_flags"], so["link_flags"])
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
        eng.step(10, dt)  # a finished simulation does not move
        assert eng.results()[0].steps == ro.steps
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


def test_force_field_and_attach_conditions():
    spec = cube_spec((3, 2, 2), seed=41, actuated=False, lift=1, name="ff")
    spec.set_program(abi.PROG_FORCE_X, ("MUL", ("CONST", 1e-3), ("SIN", ("MUL", ("VAR", "t"), ("CONST", 300.0)))))
    spec.set_program(abi.PROG_FORCE_Z, ("MUL", ("CONST", -2e-2), ("VAR", "z")))
    spec.set_program(abi.PROG_ATTACH_0, ("SUB", ("VAR", "x"), ("CONST", 0.011)))
    run_pair(spec, 800, check_every=200)


def collide_spec(sticky, detach=False, name="pile"):
    """Two 2x2x2 blocks, the upper one offset and dropped onto the lower: collisions (and sticky attach)."""
    spec = ModelSpec(0.01, name)
    if detach:
        spec.add_material(name="S", mat_model=1, elastic_mod=1e6, fail_stress=2.5e3, density=1e3, u_static=1.0, u_dynamic=0.8, sticky=int(sticky))
    else:
        spec.add_material(name="S", elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.8, sticky=int(sticky))
    spec.add_material(name="T", elastic_mod=2e6, density=1.2e3, u_static=1.0, u_dynamic=0.8, is_target=1)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01)
    spec.set_options(enable_collision=1, enable_attach=int(sticky), enable_detach=int(detach), safety_guard=50)
    st = np.zeros((5, 3, 4), np.uint8)
    st[0:2, 0:2, 0:2] = 1
    st[3:5, 1:3, 1:3] = 1
    st[0, 0, 3] = 2
    st[3, 0, 3] = 1
    spec.set_structure(st)
    return spec


@pytest.mark.parametrize("sticky", [False, True])
def test_collisions_and_attach(sticky):
    spec = collide_spec(sticky)
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        total = 0
        for _ in range(16):
            eng.step(250, dt)
            orc.step(250, dt)
            total += 250
            se, so = eng.state(0), orc.state()
            assert se["link_vneg"].shape == so["link_vneg"].shape, "link count differs at step %d" % total
            for k in ("link_vneg", "link_vpos", "link_axis", "vox_links", "link_flags", "vox_flags"):
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s at step %d" % (k, total))
            check_state(se, so, "pile step %d" % total)
            np.testing.assert_allclose(se["contact_force"], so["contact_force"], rtol=1e-9, atol=1e-18)
        re, ro = eng.results()[0], orc.result()
        assert re.collision_count == ro.collision_count
        assert re.num_links == ro.num_links
        c = orc.counts()
        if sticky:
            assert c["attach"] > 0, "scenario must produce attach events"
        else:
            assert c["attach"] == 0
        assert ro.collision_count > 0
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


def test_config4_small_pile():
    """Config 4 at oracle-checkable size: 2x2x2 grid of 3^3 sticky actuated bodies dropped onto each other
    (collisions + attach + detach enabled), hashed-grid contacts vs the oracle's all-pairs sweep."""
    from voxcraft_sim_b200 import workloads as W
    spec = W.c4_spec(grid=(2, 2, 2), body=3, name="c4small")
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        total = 0
        for i in range(8):
            eng.step(500)
            orc.step(500, -1.0)
            total += 500
            se, so = eng.state(0, link_cap=4096), orc.state()
            assert se["link_vneg"].shape == so["link_vneg"].shape, "link count differs at step %d" % total
            for k in ("link_vneg", "link_vpos", "link_axis", "vox_links", "link_flags", "vox_flags"):
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s at step %d" % (k, total))
            check_state(se, so, "c4small step %d" % total, kin=1e-8, link=1e-6)
        assert orc.counts()["attach"] > 0
    finally:
        lib.vx3_builder_destroy(b)


def test_detach():
    spec = collide_spec(True, detach=True, name="detach")
    lib, b, d = build(spec)
    d.contents.link_capacity = 4096  # repeated attach -> fail -> detach cycles keep appending to the pool
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        for i in range(10):
            eng.step(300, dt)
            orc.step(300, dt)
            se, so = eng.state(0, link_cap=4096), orc.state()
            for k in ("link_vneg", "link_vpos", "link_axis", "vox_links", "link_flags"):
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s chunk %d" % (k, i))
            compare_states(se, so, KIN, 1e-7, "detach chunk %d" % i)  # ~800 attach/detach events: chaotic, looser
    finally:
        lib.vx3_builder_destroy(b)


def test_secondary_experiment_removal_and_reinit():
    """SecondaryExperiment: voxels of a material leave the simulation after a set time (with their links), and the
    initial positions / initial CoM are re-initialised once (VX3_VoxelyzeKernel.cu:336-399)."""
    spec = cube_spec((4, 3, 3), seed=29, actuated=True, name="secondary")
    spec.materials[1]["remove_after_s"] = 0.008
    spec.set_options(secondary_experiment=1, reinit_initial_position_after_s=0.004)
    spec.set_env(temp_period=0.002)  # CoM sampled often, so the re-initialised CoM is not the initial one
    lib, b, d = build(spec)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        for i in range(6):
            eng.step(100)
            orc.step(100, -1.0)
            se, so = eng.state(0), orc.state()
            for k in ("vox_links", "link_flags", "vox_flags"):
                np.testing.assert_array_equal(se[k], so[k], err_msg="%s chunk %d" % (k, i))
            check_state(se, so, "secondary chunk %d" % i)
            re, ro = eng.results()[0], orc.result()
            np.testing.assert_allclose(list(re.initial_com), list(ro.initial_com), rtol=1e-12, atol=1e-18)
            np.testing.assert_allclose(re.total_distance_of_all_voxels, ro.total_distance_of_all_voxels, rtol=1e-9, atol=1e-15)
        assert (so["link_flags"] & abi.LINKSTATE_REMOVED).any(), "scenario must remove links"
        assert list(ro.initial_com) != [0.0, 0.0, 0.0]
    finally:
        lib.vx3_builder_destroy(b)


def test_no_device_side_cpu_fallback_symbols():
    """The product library must not contain the oracle: its only physics lives in CUDA kernels."""
    lib = util.load_engine()
    assert not hasattr(lib, "vx3o_step")
