"""GPU parity for EnableSignals / EnableCilia: the parallel replay of the sequential signal order (k_signals) and the
signal-driven cilia force against the oracle.  Signal state (localSignal, localSignaldt, inactiveUntil,
packmakerNextPulse, d_signal.value/activeTime) is bit-exact: it is pure fp64 multiply/add/compare arithmetic on times and
constants, no libm call."""
import numpy as np
import pytest

import util
from util import KIN, EngineBatch, OracleSim, compare_states
from voxcraft_sim_b200.model import ModelSpec
from test_signals import bar_spec

pytestmark = pytest.mark.gpu


def run_signal_pair(spec, chunks, chunk_steps, kin_tol=1e-9, link_cap=None):
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        seen_signal = False
        for c in range(chunks):
            eng.step(chunk_steps, dt)
            assert orc.step(chunk_steps, dt) == chunk_steps
            se, so = eng.state(0, link_cap=link_cap), orc.state()
            np.testing.assert_array_equal(se["signal"], so["signal"], err_msg="signal state, chunk %d" % c)
            compare_states(se, so, KIN, kin_tol, "chunk %d" % c)
            seen_signal |= bool(so["signal"][:, 0].any())
        assert seen_signal, "scenario must produce signals"
        return eng, orc
    finally:
        lib.vx3_builder_destroy(b)


def test_pacemaker_bar():
    run_signal_pair(bar_spec(8), chunks=40, chunk_steps=7)


def body_spec(shape=(5, 4, 3), all_pacemakers=False, delay=0.004, cilia=False, seed=3, name="sig"):
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


def test_waves_in_a_body_with_several_pacemakers():
    run_signal_pair(body_spec(), chunks=40, chunk_steps=31)


def test_all_pacemakers_long_dependency_chains():
    """Every voxel fires at once, so at the next step every voxel is an active sender and whether voxel i still sends
    depends on its lower-index neighbours' sends: the fixpoint loop of k_signals has to run many rounds."""
    run_signal_pair(body_spec((6, 3, 2), all_pacemakers=True, name="allpm"), chunks=60, chunk_steps=13)


def test_zero_time_delay_chains_within_one_step():
    """signalTimeDelay = 0: a received signal is active in the same step, so a wave can run through ascending voxel
    indices within ONE step in the sequential order; the parallel replay must reproduce it."""
    run_signal_pair(body_spec((6, 2, 2), delay=0.0, name="zerodelay"), chunks=60, chunk_steps=13)


def test_cilia_force_follows_local_signal():
    eng, orc = run_signal_pair(body_spec((4, 4, 2), cilia=True, name="cilia"), chunks=20, chunk_steps=53)


def test_target_contact_fires_signal():
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
    eng, orc = run_signal_pair(spec, chunks=30, chunk_steps=100, kin_tol=1e-8)
    assert eng.results()[0].collision_count == orc.result().collision_count > 0
