"""The time step a run with dt < 0 uses (host library, vx3_model_first_step_dt) against the oracle — which is pinned bit for bit on the
reference's doTimeStep (tests/test_oracle_vs_vx3ref.py): the reference evaluates OptimalDt once, in its first doTimeStep(dt < 0),
AFTER that step's updateTemperature (VX3_VoxelyzeKernel.cu:240-247), so with nu != 0 (stiffness eHat * area / ((1 + strain) * rest
length), VX3_Link.cu:268-277) it sees the rest lengths at the t = 0 temperatures, not the model's."""
import numpy as np
import pytest

import util
from scenarios import random_spec, random_spec2
from util import OracleSim


@pytest.mark.parametrize("family,seed", [("a", k) for k in range(40)] + [("b", k) for k in range(12)])
def test_first_step_dt_equals_the_oracles_first_step(family, seed):
    spec = (random_spec if family == "a" else random_spec2)(seed)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        first, model = lib.vx3_model_first_step_dt(d), lib.vx3_model_recommended_dt(d)
        nu = [d.contents.link_mats[d.contents.link_mat[i]].m.nu for i in range(d.contents.n_links)]
        if not any(v != 0 for v in nu):
            assert first == model  # nothing state-dependent: a1 of the link materials
        orc = OracleSim(d)
        assert orc.step(1, -1.0) == 1
        od = max(first, 1e-10)
        expect = float(np.float32(d.contents.opt.dt_frac * od))  # dt = DtFrac * OptimalDt, a float (doTimeStep(float dt))
        assert orc.result().current_time == expect, (first, model, orc.result().current_time)
    finally:
        lib.vx3_builder_destroy(b)


def test_state_dependence_is_exercised():
    """At least some of the seeds above have nu != 0 with actuation, where the two functions differ."""
    lib = util.load_engine()
    differ = 0
    for k in range(40):
        b, d = random_spec(k).build(lib)
        differ += lib.vx3_model_first_step_dt(d) != lib.vx3_model_recommended_dt(d)
        lib.vx3_builder_destroy(b)
    assert differ >= 3, differ
