"""GPU checks at BASELINE.json's sizes, through properties that do not need the (slow) oracle for the whole run:
  * config 2 at full size: the on-chip persistent kernel (148 blocks, face links evaluated on both sides, one exchange per
    step) and the streaming kernels agree bit for bit, and both match the oracle over a short horizon;
  * config 3 (sample of the batch): a simulation steps identically inside a batch and alone, the two link-pass variants
    are bit-identical, and sampled robots match the oracle;
  * config 5 (reduced length, full cross-section): link-pass variants bit-identical on a single large body."""
import os

import numpy as np
import pytest

import util
from util import KIN, LINKF, LINKS, EngineBatch, OracleSim, compare_states
from voxcraft_sim_b200 import workloads as W

pytestmark = pytest.mark.gpu

BITS = KIN + LINKF + LINKS + ["link_flags", "vox_flags", "temp", "link_rest_length", "link_strain", "link_max_strain", "link_stress"]


def test_config2_full_size_persistent_equals_streaming_and_oracle():
    spec = W.c2_spec()
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        assert d.contents.n_voxels == 8000 and d.contents.n_links == 22800
        states = []
        for use in (True, False):
            eng = EngineBatch([d])
            eng.set_profiling(False, use_persistent=use)
            eng.step(150)
            s150 = eng.state(0)
            eng.step(450)  # 600 steps in all: crosses launches of the persistent kernel
            states.append((s150, eng.state(0), eng.results()[0]))
            eng.close()
        util.assert_bit_equal(states[0][0], states[1][0], BITS, "config 2, 150 steps, persistent vs streaming")
        util.assert_bit_equal(states[0][1], states[1][1], BITS, "config 2, 600 steps, persistent vs streaming")
        assert states[0][2].steps == states[1][2].steps == 600
        assert states[0][2].current_time == states[1][2].current_time
        orc = OracleSim(d)
        assert orc.step(150, -1.0) == 150
        so = orc.state()
        # positions / orientations at SURVEY 8(d)'s 1e-9; momenta are small differences of large per-step force sums, so the
        # 1-ulp libdevice-vs-glibc sin() seed shows ~1e-8 there on an 8000-voxel body (measured 2.6e-8): gate 1e-6
        compare_states(states[0][0], so, ["pos", "orient"], 1e-9, "config 2 vs oracle, 150 steps")
        compare_states(states[0][0], so, ["lin_mom", "ang_mom"], 1e-6, "config 2 vs oracle, 150 steps")
        compare_states(states[0][0], so, LINKF + LINKS, 1e-6, "config 2 vs oracle, 150 steps")
        np.testing.assert_array_equal(states[0][0]["link_flags"], so["link_flags"])
        np.testing.assert_array_equal(states[0][0]["vox_flags"], so["vox_flags"])
    finally:
        lib.vx3_builder_destroy(b)


def _with_env(name, value, fn):
    old = os.environ.get(name)
    try:
        if value is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = value
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


def test_config3_sample_batch_properties():
    lib = util.load_engine()
    built = [W.c3_spec(k).build(lib) for k in range(48)]
    descs = [d for _, d in built]
    try:
        def run(variant, which=None, steps=400):
            def go():
                eng = EngineBatch(descs if which is None else [descs[which]])
                eng.step(steps)
                out = [eng.state(i) for i in ([3, 17, 40] if which is None else [0])]
                eng.close()
                return out
            return _with_env("VX3_LINK_QUEUE", variant, go)
        base = run("0")
        for variant in ("1", None):  # deferred dense passes, timing-based choice
            other = run(variant)
            for a, b_ in zip(base, other):
                util.assert_bit_equal(a, b_, BITS, "config 3 sample, link pass variant %s" % variant)
        # a robot steps the same inside the batch and alone (alone, it takes the on-chip persistent path)
        for j, k in enumerate([3, 17, 40]):
            alone = run(None, which=k)[0]
            util.assert_bit_equal(base[j], alone, BITS, "config 3 robot %d: batch vs alone" % k)
        # ... and like the oracle
        for j, k in ((0, 3), (2, 40)):
            orc = OracleSim(descs[k])
            assert orc.step(400, -1.0) == 400
            so = orc.state()
            compare_states(base[j], so, ["pos", "orient"], 1e-9, "config 3 robot %d vs oracle" % k)
            compare_states(base[j], so, ["lin_mom", "ang_mom"], 1e-6, "config 3 robot %d vs oracle" % k)
            np.testing.assert_array_equal(base[j]["link_flags"], so["link_flags"])
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)


def test_config5_cross_section_link_variants_bit_identical():
    spec = W.c5_spec((24, 200, 100))  # the full 200 x 100 face, 24 voxels long (480k voxels, 1.4M links)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        def run(variant):
            def go():
                eng = EngineBatch([d])
                eng.step(40)
                st = eng.state(0)
                r = eng.results()[0]
                eng.close()
                return st, r
            return _with_env("VX3_LINK_QUEUE", variant, go)
        s0, r0 = run("0")
        for variant in ("1",):
            s1, r1 = run(variant)
            util.assert_bit_equal(s0, s1, BITS, "config 5 slice, link pass variant %s" % variant)
            assert list(r0.current_com) == list(r1.current_com)
        assert r0.status == 0 and np.isfinite(s0["pos"]).all()
    finally:
        lib.vx3_builder_destroy(b)


def test_config3_full_run_fitness_matches_the_oracle():
    """SURVEY §8(d), last gate: a config 3 robot run to its stop condition (t > 1 s, ~40,000 steps, floor contact and
    stick/slip friction included) reports the oracle's step count and time exactly and its fitness within 1e-6 relative."""
    lib = util.load_engine()
    picks = [W.c3_spec(k) for k in (5, 12)]
    built = [s.build(lib) for s in picks]
    try:
        eng = EngineBatch([d for _, d in built])
        eng.run()
        res = eng.results()
        eng.close()
        for (b, d), r in zip(built, res):
            orc = OracleSim(d)
            orc.run()
            ro = orc.result(refresh=False)
            rel = abs(r.fitness_score - ro.fitness_score) / max(abs(ro.fitness_score), 1e-30)
            print("config 3 full run: %d voxels, %d steps, fitness %.9g (oracle %.9g, rel %.1e)" % (d.contents.n_voxels, r.steps, r.fitness_score, ro.fitness_score, rel))
            assert r.status == ro.status == 1  # VX3_SIM_STOPPED
            assert r.steps == ro.steps and r.current_time == ro.current_time
            np.testing.assert_allclose(r.fitness_score, ro.fitness_score, rtol=1e-6)
            np.testing.assert_allclose(list(r.current_com), list(ro.current_com), rtol=1e-6)
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)
