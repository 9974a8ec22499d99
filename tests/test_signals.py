"""Signals (EnableSignals; VX3_Voxel.cu:279-348): known-answer checks of the oracle's restatement derived by hand from
the reference's receiveSignal / propagateSignal / packMaker / localSignalDecay, on a 1-D bar whose first voxel is a
pacemaker.  CPU only; the GPU twin is tests/test_gpu_signals.py."""
import numpy as np

import util
from util import OracleSim
from voxcraft_sim_b200.model import ModelSpec


def bar_spec(n=8, period=0.05, delay=0.01, inactive=0.03, decay=0.9, name="bar"):
    spec = ModelSpec(0.01, name)
    common = dict(elastic_mod=1e6, density=1e3, u_static=1.0, u_dynamic=0.8, signal_time_delay=delay, inactive_period=inactive,
                  signal_value_decay=decay)
    spec.add_material(name="P", is_pacemaker=1, pacemaker_period=period, **common)
    spec.add_material(name="N", **common)
    spec.set_env(bond_damping_z=1.0, col_damping_z=0.8, slow_damping_z=0.01)
    spec.set_options(enable_collision=0, enable_signals=1)
    st = np.full((1, 1, n), 2, np.uint8)
    st[0, 0, 0] = 1
    spec.set_structure(st)
    return spec


def test_pacemaker_wave_known_answer():
    n, period, delay, inactive, decay = 8, 0.05, 0.01, 0.03, 0.9
    spec = bar_spec(n, period, delay, inactive, decay)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        # ---- hand model of the first pulse travelling down the bar (times accumulate like currentTime += dt) ----
        # step 0 (t=0): packMaker fires on voxel 0: localSignal=100, d_signal={100*decay, 0}; decay the same step -> 90
        # step s with t_s >= activeTime of voxel k: voxel k sends value_k to k+1, which becomes active at t_s + delay
        nsteps = int(0.12 / dt)
        t, times = 0.0, []
        for _ in range(nsteps + 1):
            times.append(t)
            t += dt
        arrive_step = {0: 0}
        active = {0: 0.0}
        value = {0: 100.0}
        for k in range(1, n):
            s = next(i for i in range(arrive_step[k - 1] + 1, nsteps) if times[i] >= active[k - 1])
            arrive_step[k] = s
            active[k] = times[s] + delay
            value[k] = value[k - 1] * decay  # localSignal of k = d_signal.value of k-1
        seen = {}
        for s in range(arrive_step[n - 1] + 2):
            orc.step(1, dt)
            sig = orc.state()["signal"]
            for k in range(n):
                if k not in seen and sig[k, 0] > 0:
                    seen[k] = (s, sig[k, 0])
        for k in range(n):
            assert seen[k][0] == arrive_step[k], "voxel %d: wave arrived at step %d, expected %d" % (k, seen[k][0], arrive_step[k])
            # on arrival localSignal is the received value, decayed once in the same step when localSignaldt <= t
            # (always true on first arrival: localSignaldt starts at 0) — voxel 0 likewise (packMaker then decay)
            np.testing.assert_allclose(seen[k][1], value[k] * 0.9, rtol=1e-15)
        # packmakerNextPulse: the second pulse fired at the first step with t >= period
        sig = orc.state()["signal"]
        t2 = next(x for x in times if x >= period)
        assert arrive_step[n - 1] + 1 >= times.index(t2), "run long enough to see the second pulse"
        np.testing.assert_allclose(sig[0, 3], t2 + period, rtol=1e-15)
    finally:
        lib.vx3_builder_destroy(b)


def test_pacemaker_repeats_and_local_signal_decays():
    spec = bar_spec(4, period=0.02, delay=0.004, inactive=0.004)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc = OracleSim(d)
        dt = float(np.float32(0.9 * orc.recommended_dt()))
        pulses, prev = 0, 0.0
        ls0 = []
        t = 0.0
        for s in range(int(0.07 / dt)):
            orc.step(1, dt)
            sig = orc.state()["signal"]
            ls0.append(sig[0, 0])
            if sig[0, 3] != prev:  # packmakerNextPulse moved: a pulse fired at this step
                pulses += 1
                np.testing.assert_allclose(sig[0, 3], t + 0.02, rtol=1e-15)
                prev = sig[0, 3]
            t += dt
        assert pulses == 4  # t = 0, then the first step at or after each multiple of the period
        ls0 = np.array(ls0)
        # between pulses localSignal only falls, by the factor 0.9 every 0.01 s (localSignalDecay)
        drops = ls0[1:] / np.maximum(ls0[:-1], 1e-300)
        assert set(np.round(drops[(drops < 1.0) & (ls0[:-1] > 0)], 12)) <= {0.9}
    finally:
        lib.vx3_builder_destroy(b)


def test_signals_off_leaves_state_zero():
    spec = bar_spec(4)
    spec.set_options(enable_signals=0)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc = OracleSim(d)
        orc.step(50, -1.0)
        assert not orc.state()["signal"].any()
    finally:
        lib.vx3_builder_destroy(b)
