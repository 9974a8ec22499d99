"""Named, seeded scenarios shared by the CPU pin (oracle vs the reference's own VX3 code, tests/test_oracle_vs_vx3ref.py),
the golden-fixture generator (tests/golden/make_golden_vx3.py) and the GPU parity tests.  Each entry:

    name -> dict(spec=callable -> ModelSpec, steps=total doTimeStep calls, chunk=compare every `chunk` steps,
                 dt="fixed" (0.9 * recommendedTimeStep passed explicitly) | "auto" (doTimeStep(-1)),
                 link_capacity=pool size override or None, covers="what reference code the scenario drives")
"""
import numpy as np

from voxcraft_sim_b200 import abi
from voxcraft_sim_b200 import workloads as W
from voxcraft_sim_b200.model import ModelSpec
from voxcraft_sim_b200.workloads import add_abc_materials, splitmix64


def cube_spec(n=(3, 3, 3), seed=42, actuated=True, lift=0, holes=0.0, name="cube", collisions=0, damping=(1.0, 0.8, 0.01)):
    """Multi-material (A/B/C) body of nx*ny*nz voxels, `lift` empty layers below, optional random holes."""
    nx, ny, nz = n
    spec = ModelSpec(0.01, name)
    add_abc_materials(spec)
    spec.set_env(bond_damping_z=damping[0], col_damping_z=damping[1], slow_damping_z=damping[2],
                 temp_enabled=1, vary_temp_enabled=int(actuated), temp_amplitude=20.0 if actuated else 0.0, temp_period=0.2)
    spec.set_options(enable_collision=collisions)
    r = splitmix64(seed)
    r2 = splitmix64(seed + 1)
    r3 = splitmix64(seed + 2)
    st = np.zeros((nz + lift, ny, nx), np.uint8)
    ph = np.zeros((nz + lift, ny, nx))
    for z in range(nz):
        for y in range(ny):
            for x in range(nx):
                m = 1 + r() % 3
                p = (r2() >> 11) / float(1 << 53)
                keep = ((r3() >> 11) / float(1 << 53)) >= holes
                if keep:
                    st[z + lift, y, x] = m
                ph[z + lift, y, x] = p
    spec.set_structure(st, phase_offset=ph if actuated else None)
    return spec


def collide_spec(sticky, detach=False, name="pile", nu=0.0):
    """Two 2x2x2 blocks, the upper one offset and dropped onto the lower: collisions (and sticky attach)."""
    spec = ModelSpec(0.01, name)
    if detach:
        spec.add_material(name="S", mat_model=1, elastic_mod=1e6, fail_stress=2.5e3, density=1e3, u_static=1.0, u_dynamic=0.8, sticky=int(sticky))
    else:
        spec.add_material(name="S", elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.8, sticky=int(sticky), poissons_ratio=nu)
    spec.add_material(name="T", elastic_mod=2e6, density=1.2e3, u_static=1.0, u_dynamic=0.8, is_target=1)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01, volume_effects_enabled=int(nu != 0))  # (nu is forced to 0 without it)
    if nu != 0:  # actuation, so that the voxels are strained when the attach makes its link (the new link's transverse info takes that in)
        spec.set_env(temp_enabled=1, vary_temp_enabled=1, temp_amplitude=15.0, temp_period=0.02)
        for m in spec.materials:
            m["cte"] = 0.01
    spec.set_options(enable_collision=1, enable_attach=int(sticky), enable_detach=int(detach), safety_guard=50)
    st = np.zeros((5, 3, 4), np.uint8)
    st[0:2, 0:2, 0:2] = 1
    st[3:5, 1:3, 1:3] = 1
    st[0, 0, 3] = 2
    st[3, 0, 3] = 1
    spec.set_structure(st)
    return spec


def cantilever_spec():
    """A passive cantilever with a fixed root and an end load: drives links out of the small-angle regime."""
    spec = ModelSpec(0.01, "cantilever")
    spec.add_material(elastic_mod=2e5, density=1e3, u_static=1.0, u_dynamic=0.5)
    spec.set_env(bond_damping_z=0.5, col_damping_z=0.8, slow_damping_z=0.02, floor_enabled=0)
    spec.set_structure(np.ones((1, 1, 10), np.uint8))
    spec.set_external(0, dof_fixed=0x3F)
    spec.set_external(9, force=(0.0, 0.0, -0.02))
    return spec


def forcefield_spec():
    spec = cube_spec((3, 2, 2), seed=41, actuated=False, lift=1, name="ff")
    spec.set_program(abi.PROG_FORCE_X, ("MUL", ("CONST", 1e-3), ("SIN", ("MUL", ("VAR", "t"), ("CONST", 300.0)))))
    spec.set_program(abi.PROG_FORCE_Z, ("MUL", ("CONST", -2e-2), ("VAR", "z")))
    spec.set_program(abi.PROG_ATTACH_0, ("SUB", ("VAR", "x"), ("CONST", 0.011)))
    return spec


def secondary_spec():
    spec = cube_spec((4, 3, 3), seed=29, actuated=True, name="secondary")
    spec.materials[1]["remove_after_s"] = 0.008
    spec.set_options(secondary_experiment=1, reinit_initial_position_after_s=0.004)
    spec.set_env(temp_period=0.002)  # CoM sampled often, so the re-initialised CoM is not the initial one
    return spec


def runner_spec():
    spec = cube_spec((3, 3, 2), seed=31, actuated=True, name="runner")
    spec.set_env(temp_period=0.01)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.05)))
    spec.set_program(abi.PROG_FITNESS, ("SQRT", ("ADD", ("MUL", ("VAR", "x"), ("VAR", "x")), ("MUL", ("VAR", "y"), ("VAR", "y")))))
    return spec


def history_spec():
    """runner_spec with the history recorder on: voxel frames and link frames, the CUDA_Simulation stdout is the fixture."""
    spec = runner_spec()
    spec.name = "runner.vxd"
    spec.set_options(record_step_size=200, record_voxel=1, record_link=1)
    return spec


def signal_body_spec(shape=(5, 4, 3), all_pacemakers=False, delay=0.004, cilia=False, seed=3, name="sig"):
    spec = ModelSpec(0.01, name)
    common = dict(elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.8, inactive_period=0.006)
    spec.add_material(name="P", is_pacemaker=1, pacemaker_period=0.013, signal_time_delay=delay, cilia=1.0 if cilia else 0.0, **common)
    spec.add_material(name="N", is_pacemaker=int(all_pacemakers), pacemaker_period=0.017, signal_time_delay=delay * 1.5,
                      signal_value_decay=0.8, cilia=1.0 if cilia else 0.0, **common)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01)
    spec.set_options(enable_collision=0, enable_signals=1, enable_cilia=int(cilia))
    nx, ny, nz = shape
    rng = np.random.RandomState(seed)
    st = np.full((nz, ny, nx), 2, np.uint8)
    st[rng.rand(nz, ny, nx) < 0.08] = 1
    st[0, 0, 0] = 1
    st[rng.rand(nz, ny, nx) < 0.1] = 0  # holes: irregular neighbourhoods
    st[0, 0, 0] = 1
    kw = {}
    if cilia:
        kw["base_cilia"] = rng.uniform(-1e-4, 1e-4, (nz, ny, nx, 3))
        kw["shift_cilia"] = rng.uniform(-1e-6, 1e-6, (nz, ny, nx, 3))
    spec.set_structure(st, **kw)
    return spec


def touch_spec():
    """A non-target voxel touching a target voxel receives a forced signal (VX3_VoxelyzeKernel.cu:719-725)."""
    spec = ModelSpec(0.01, "touch")
    spec.add_material(name="S", elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.8, signal_time_delay=0.002, inactive_period=0.004)
    spec.add_material(name="T", elastic_mod=2e6, density=1.2e3, u_static=1.0, u_dynamic=0.8, is_target=1)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01)
    spec.set_options(enable_collision=1, enable_signals=1)
    st = np.zeros((5, 2, 3), np.uint8)
    st[0:2, 0:2, 0:3] = 1   # lower block
    st[3:5, 0:2, 0:2] = 2   # target block dropped onto it
    spec.set_structure(st)
    return spec


def closeness_spec():
    """Target voxels + MaxDistInVoxelLengthsToCountAsPair: computeTargetCloseness / numClosePairs / recentAngle at the
    CoM sampling cadence (VX3_VoxelyzeKernel.cu:314-334, 545-563) feeding a fitness that reads them."""
    spec = cube_spec((4, 3, 2), seed=57, actuated=True, name="closeness")
    spec.materials[2]["is_target"] = 1
    spec.set_env(temp_period=0.004)
    spec.set_options(max_dist_in_voxel_lengths_to_count_as_pair=2.5)
    spec.set_program(abi.PROG_FITNESS, ("ADD", ("ADD", ("VAR", "targetCloseness"), ("VAR", "numClosePairs")),
                                        ("ADD", ("VAR", "angle"), ("MUL", ("VAR", "num_voxel"), ("VAR", "hit")))))
    return spec


# ---------------------------------------------------------------- nonlinear materials / Poisson's ratio
def bar_material_spec(mat_model, nu=0.0, name="bar", fail=False, n=8, force=2.2):
    """A bar fixed at one end and pulled at the other by a constant force, lightly damped: the first overshoot drives the
    links past the yield point (loading branch, VX3_Link.cu:220-245), the rebound unloads them along the elastic slope
    from the plastic offset (:246-256), the following swings re-load — and, with `fail`, a link passes the failure
    strain so its force drops to zero (VX3_Material.cu:90-95, VX3_Link.cu:178-184).  mat_model: 1 linear+fail,
    2 bilinear, 3 piece-wise data (src/VXA/VX_Object.cpp:1395-1460).  nu != 0 needs VolumeEffectsEnabled
    (src/VXA/VX_Sim.cpp:148)."""
    spec = ModelSpec(0.01, name)
    kw = dict(density=1e3, u_static=1.0, u_dynamic=0.5, poissons_ratio=nu)
    E = 1e6
    if mat_model == 0:
        spec.add_material(mat_model=0, elastic_mod=E, **kw)
    elif mat_model == 1:
        spec.add_material(mat_model=1, elastic_mod=E, fail_stress=3.5e4 if fail else 1e9, **kw)
    elif mat_model == 2:
        spec.add_material(mat_model=2, elastic_mod=E, plastic_mod=2e5, yield_stress=1.2e4, fail_stress=2.4e4 if fail else 1e9, **kw)
    else:
        strain = [0.0, 0.01, 0.03, 0.08, 0.2]
        stress = [0.0, 1e4, 1.6e4, 2.2e4, 2.6e4 if fail else 3.5e4]
        spec.add_material(mat_model=3, elastic_mod=E, n_data=len(strain), strain_data=strain, stress_data=stress, **kw)
    spec.set_env(bond_damping_z=0.3, col_damping_z=0.8, slow_damping_z=0.002, floor_enabled=0, grav_enabled=0,
                 volume_effects_enabled=int(nu != 0.0))
    spec.set_options(enable_collision=0)
    st = np.ones((2, 2, n), np.uint8)
    spec.set_structure(st)
    for k in range(4):
        spec.set_external(n * k, dof_fixed=0x3F)
        spec.set_external(n * k + n - 1, force=(force * (1.0 + 0.1 * k), 0.02 * k, -0.03))
    return spec


def random_spec(seed):
    """A seeded random model over the option space of the step loop: shape, holes, a 2-3 entry palette with random stiffness /
    density / CTE / friction / Poisson's ratio / material model, actuation with random per-voxel phases (contraction-only or with
    expansion), floor on or off, a thermal start delay, an external force or a fixed voxel, and — on every second seed —
    collisions with sticky attach (and detach on every fourth).  Parameters stay in the range where the run is stable."""
    rs = np.random.RandomState(1000 + seed)
    U = lambda a, b: float(rs.uniform(a, b))
    nx, ny, nz = int(rs.randint(2, 5)), int(rs.randint(2, 5)), int(rs.randint(2, 4))
    collide = seed % 2 == 1
    detach = seed % 4 == 3
    spec = ModelSpec(0.01, "fuzz%d" % seed)
    npal = int(rs.randint(2, 4))
    for k in range(npal):
        model = int(rs.choice([0, 0, 1, 2])) if not detach else 1
        kw = dict(name="M%d" % k, elastic_mod=U(3e5, 2e6), density=U(800, 1500), cte=float(rs.choice([-1, 1])) * U(0.004, 0.015),
                  u_static=U(0.6, 1.2), u_dynamic=U(0.3, 0.6), poissons_ratio=float(rs.choice([0.0, 0.0, U(0.15, 0.35)])),
                  material_temp_phase=U(0, 1), thermal_on_after_s=float(rs.choice([0.0, U(0.0005, 0.003)])))
        if model == 1:
            kw.update(mat_model=1, fail_stress=kw["elastic_mod"] * (U(0.1, 0.2) if detach else 0.5))
        elif model == 2:
            ys = kw["elastic_mod"] * U(0.01, 0.03)
            kw.update(mat_model=2, plastic_mod=kw["elastic_mod"] * U(0.1, 0.5), yield_stress=ys, fail_stress=ys * U(3, 6))
        if collide:
            kw.update(sticky=1)
        spec.add_material(**kw)
    if collide:  # every voxel of a sticky pile is the same material: attach needs equal materials on both sides
        spec.materials = spec.materials[:1]
        npal = 1
    spec.set_env(bond_damping_z=U(0.5, 1.0), col_damping_z=U(0.5, 0.9), slow_damping_z=U(0.005, 0.03), floor_enabled=int(rs.rand() < 0.8),
                 temp_enabled=1, vary_temp_enabled=1, temp_amplitude=U(2, 6) if collide else U(8, 25), temp_period=U(0.01, 0.05),
                 volume_effects_enabled=int(rs.rand() < 0.3))
    spec.set_options(enable_collision=int(collide), enable_attach=int(collide), enable_detach=int(detach), safety_guard=int(rs.randint(20, 80)),
                     enable_expansion=int(rs.rand() < 0.4))
    lift = int(rs.randint(0, 3))
    if collide:  # two bodies one empty layer apart, the upper one shifted by a cell; strong gravity so that they meet within a few hundred steps
        spec.set_env(grav_acc=-U(150.0, 400.0))
        st = np.zeros((2 * nz + 1 + lift, ny + 1, nx + 1), np.uint8)
        st[lift:lift + nz, :ny, :nx] = 1
        st[lift + nz + 1:lift + 2 * nz + 1, 1:, 1:] = 1
    else:
        st = np.zeros((nz + lift, ny, nx), np.uint8)
        st[lift:] = rs.randint(1, npal + 1, size=(nz, ny, nx))
        st[lift:][rs.rand(nz, ny, nx) < U(0.0, 0.25)] = 0
        if not st.any():
            st[lift, 0, 0] = 1
    spec.set_structure(st, phase_offset=rs.rand(*st.shape))
    nvox = int((st > 0).sum())
    if not collide and nvox > 2:
        if rs.rand() < 0.5:
            spec.set_external(int(rs.randint(0, nvox)), force=(U(-2e-3, 2e-3), U(-2e-3, 2e-3), U(-2e-3, 2e-3)))
        if rs.rand() < 0.4:
            spec.set_external(0, dof_fixed=int(rs.choice([0x3F, 0x07, 0x04])))
    return spec


def random_spec2(seed):
    """Second random family: the OPTIONAL physics in random combination — signals (pacemaker materials, random delays / decay /
    inactive periods), cilia driven by signals, SecondaryExperiment (a material that is removed, CoM re-initialisation), force-field
    and attach-condition programs, target materials with pair counting, a fixed material, collisions with a dropped body on every
    third seed (so a touch can fire a signal)."""
    rs = np.random.RandomState(5000 + seed)
    U = lambda a, b: float(rs.uniform(a, b))
    nx, ny, nz = int(rs.randint(3, 6)), int(rs.randint(2, 5)), int(rs.randint(2, 4))
    signals, cilia = bool(rs.rand() < 0.7), bool(rs.rand() < 0.5)
    secondary, collide = bool(rs.rand() < 0.4), seed % 3 == 2
    spec = ModelSpec(0.01, "fuzzb%d" % seed)
    npal = 3
    for k in range(npal):
        spec.add_material(name="Q%d" % k, elastic_mod=U(4e5, 1.5e6), density=U(900, 1400), cte=float(rs.choice([-1, 1])) * U(0.003, 0.012),
                          u_static=U(0.6, 1.2), u_dynamic=U(0.3, 0.6), material_temp_phase=U(0, 1),
                          is_pacemaker=int(signals and (k == 0 or rs.rand() < 0.3)), pacemaker_period=U(0.002, 0.006),
                          signal_time_delay=float(rs.choice([0.0, U(0.0005, 0.002)])), signal_value_decay=U(0.6, 0.95),
                          inactive_period=U(0.001, 0.003), cilia=(U(0.5, 2.0) if cilia else 0.0), cilia_on_after_s=float(rs.choice([0.0, U(0.0005, 0.002)])),
                          is_target=int(k == 2 and rs.rand() < 0.6), is_measured=int(rs.rand() < 0.8),
                          remove_after_s=(U(0.002, 0.004) if secondary and k == 1 else 0.0), fixed=int(k == 2 and rs.rand() < 0.15))
    spec.set_env(bond_damping_z=U(0.5, 1.0), col_damping_z=U(0.5, 0.9), slow_damping_z=U(0.005, 0.03), temp_enabled=1, vary_temp_enabled=int(rs.rand() < 0.7),
                 temp_amplitude=U(5, 20), temp_period=U(0.002, 0.01))
    spec.set_options(enable_collision=int(collide), enable_signals=int(signals), enable_cilia=int(cilia), secondary_experiment=int(secondary),
                     reinit_initial_position_after_s=(U(0.001, 0.003) if secondary else 0.0), enable_expansion=int(rs.rand() < 0.3),
                     max_dist_in_voxel_lengths_to_count_as_pair=float(rs.choice([0.0, U(1.5, 3.0)])))
    lift = int(rs.randint(0, 2))
    extra = nz + 1 if collide else 0
    st = np.zeros((nz + lift + extra, ny, nx), np.uint8)
    st[lift:lift + nz] = rs.randint(1, npal + 1, size=(nz, ny, nx))
    st[lift:lift + nz][rs.rand(nz, ny, nx) < U(0.0, 0.2)] = 0
    st[lift, 0, 0] = 1
    if collide:
        spec.set_env(grav_acc=-U(150.0, 400.0))
        st[lift + nz + 1:lift + 2 * nz + 1, : max(ny - 1, 1), : max(nx - 1, 1)] = rs.randint(1, npal + 1, size=(nz, max(ny - 1, 1), max(nx - 1, 1)))
    kw = {}
    if cilia:
        kw["base_cilia"] = rs.uniform(-1e-4, 1e-4, st.shape + (3,))
        kw["shift_cilia"] = rs.uniform(-1e-6, 1e-6, st.shape + (3,))
    spec.set_structure(st, phase_offset=rs.rand(*st.shape), **kw)
    if rs.rand() < 0.5:
        spec.set_program(abi.PROG_FORCE_X, ("MUL", ("CONST", U(1e-4, 1e-3)), ("SIN", ("MUL", ("VAR", "t"), ("CONST", U(200.0, 2000.0))))))
        spec.set_program(abi.PROG_FORCE_Z, ("MUL", ("CONST", -U(1e-3, 2e-2)), ("VAR", "z")))
    if rs.rand() < 0.3:
        spec.set_program(abi.PROG_FITNESS, ("ADD", ("VAR", "x"), ("ADD", ("VAR", "targetCloseness"), ("MUL", ("VAR", "numClosePairs"), ("VAR", "hit")))))
    return spec


def random_spec3(seed):
    """Third random family: everything at once — random_spec2 (signals, cilia, removal, programs, targets) with strong gravity,
    every material sticky and linear-with-failure, nu != 0 on some of them with volume effects on, collisions + attach, detach on
    every second seed."""
    spec = random_spec2(seed)
    rs = np.random.RandomState(9000 + seed)
    spec.name = "fuzzc%d" % seed
    spec.set_env(volume_effects_enabled=1, grav_acc=-float(rs.uniform(150, 400)))
    for m in spec.materials:
        m["poissons_ratio"] = float(rs.choice([0.0, rs.uniform(0.15, 0.35)]))
        m["sticky"] = 1
        m["mat_model"] = 1
        m["fail_stress"] = m["elastic_mod"] * float(rs.uniform(0.1, 0.3))
    spec.set_options(enable_collision=1, enable_attach=1, enable_detach=int(seed % 2), safety_guard=int(rs.randint(20, 80)))
    return spec


SCENARIOS = {
    "act333": dict(spec=lambda: cube_spec((3, 3, 3), seed=11, actuated=True, name="act333"), steps=1000, chunk=250,
                   covers="per-voxel phase actuation, rest length from end temperatures"),
    "ragged": dict(spec=lambda: cube_spec((5, 4, 3), seed=5, actuated=True, lift=2, holes=0.25, name="ragged"), steps=1500, chunk=500,
                   covers="free fall, floor, kinetic/static friction incl. angMom=0 (VX3_Voxel.cu:259-264)"),
    "cantilever": dict(spec=cantilever_spec, steps=6000, chunk=2000, covers="fixed DOFs, external force, large-angle orientLink"),
    "forcefield": dict(spec=forcefield_spec, steps=800, chunk=200, covers="force-field programs, attach condition program"),
    "pile": dict(spec=lambda: collide_spec(False), steps=4000, chunk=1000, covers="all-pairs sweep, VX3_Collision, collisionCount"),
    "pile_sticky": dict(spec=lambda: collide_spec(True, name="pile_sticky"), steps=4000, chunk=1000,
                        covers="attach: link ctor, combinedMaterial, isNewLink ramp, surface regeneration"),
    "pile_sticky_nu": dict(spec=lambda: collide_spec(True, name="pile_sticky_nu", nu=0.3), steps=4000, chunk=1000,
                           covers="attach with nu != 0: the new link's frozen transverse area / strain sum from the end voxels' poissons strain "
                                  "(VX3_Link.cu:58-70, 85-88; VX3_Voxel.cu:428-513)"),
    "detach": dict(spec=lambda: collide_spec(True, detach=True, name="detach"), steps=3000, chunk=750, link_capacity=4096,
                   covers="gpu_update_detach, repeated attach/fail/detach cycles"),
    "c4small": dict(spec=lambda: W.c4_spec(grid=(2, 2, 2), body=3, name="c4small"), steps=2000, chunk=500, dt="auto", link_capacity=4096,
                    covers="config 4 at toy size: actuated sticky bodies, collisions + attach + detach"),
    "secondary": dict(spec=secondary_spec, steps=600, chunk=100, dt="auto", covers="SecondaryExperiment removeVoxels + reinit"),
    "sig_body": dict(spec=lambda: signal_body_spec(), steps=1240, chunk=310, covers="signals: pacemaker, propagate, decay"),
    "sig_allpm": dict(spec=lambda: signal_body_spec((6, 3, 2), all_pacemakers=True, name="allpm"), steps=780, chunk=195,
                      covers="signals: long same-step dependency chains"),
    "sig_zerodelay": dict(spec=lambda: signal_body_spec((6, 2, 2), delay=0.0, name="zerodelay"), steps=780, chunk=195,
                          covers="signals: zero time delay"),
    "cilia": dict(spec=lambda: signal_body_spec((4, 4, 2), cilia=True, name="cilia"), steps=1060, chunk=265,
                  covers="gpu_update_cilia_force with localSignal shift"),
    "touch": dict(spec=touch_spec, steps=3000, chunk=750, covers="target contact fires a forced signal"),
    "closeness": dict(spec=closeness_spec, steps=1200, chunk=300, dt="auto", covers="targetCloseness, numClosePairs, recentAngle, fitness vars"),
    "bilinear": dict(spec=lambda: bar_material_spec(2, name="bilinear"), steps=1500, chunk=250, covers="MatModel 2: yield, unload, reload"),
    "bilinear_fail": dict(spec=lambda: bar_material_spec(2, fail=True, name="bilinear_fail", force=2.6), steps=300, chunk=75,
                          covers="MatModel 2 through failure"),
    "datamat": dict(spec=lambda: bar_material_spec(3, name="datamat"), steps=1500, chunk=250, covers="MatModel 3: piece-wise lookup"),
    "datamat_fail": dict(spec=lambda: bar_material_spec(3, fail=True, name="datamat_fail", force=2.9), steps=300, chunk=75,
                         covers="MatModel 3 through the last data point (failure)"),
    "linfail": dict(spec=lambda: bar_material_spec(1, fail=True, name="linfail", force=3.8), steps=300, chunk=75,
                    covers="MatModel 1 failure (the freed end flies off: the run stops before strain 100 = divergence)"),
    "poisson_lin": dict(spec=lambda: bar_material_spec(0, nu=0.3, name="poisson_lin"), steps=1000, chunk=250,
                        covers="nu != 0, VolumeEffectsEnabled: eHat path of VX3_Material::stress"),
    "poisson_bilinear": dict(spec=lambda: bar_material_spec(2, nu=0.3, name="poisson_bilinear"), steps=1500, chunk=250,
                             covers="nu != 0 with plastic offset (VX3_Link.cu:238-241, 249-251)"),
    "poisson_data": dict(spec=lambda: bar_material_spec(3, nu=0.25, name="poisson_data"), steps=1500, chunk=250,
                         covers="nu != 0 on the piece-wise branch (VX3_Material.cu:108-121)"),
}
# seeded random models (random_spec): the same three-way check as the named scenarios — reference VX3 code = oracle bit for bit,
# GPU within the gate — on option combinations nobody picked by hand
for _k in range(8):
    SCENARIOS["fuzz%d" % _k] = dict(spec=(lambda k=_k: random_spec(k)), steps=1600 if _k % 2 else 600, chunk=400 if _k % 2 else 150, dt="auto" if _k % 2 else "fixed",
                                    link_capacity=2048 if _k % 2 else None,
                                    covers="random model %d: %s" % (_k, "two sticky bodies, collisions + attach%s" % (" + detach" if _k % 4 == 3 else "")
                                                                    if _k % 2 else "random palette / actuation / externals"))
for _k in (0, 2, 8, 11):
    SCENARIOS["fuzzb%d" % _k] = dict(spec=(lambda k=_k: random_spec2(k)), steps=1600, chunk=400, dt="auto" if _k % 2 else "fixed", link_capacity=None,
                                     covers="random model b%d: signals / cilia / removal / programs / targets in random combination" % _k)


def scenario(name):
    s = dict(SCENARIOS[name])
    s.setdefault("dt", "fixed")
    s.setdefault("link_capacity", None)
    s["name"] = name
    return s
