"""Generates tests/golden/vx3_*.json from the reference's OWN VX3 step loop (src/VX3/*.cu compiled unmodified for the host
into oracle/_ref/libvxref_vx3.so by `make -C oracle ref_vx3`, see oracle/ref_vx3/vxhost.h).  Run where the reference tree
exists:

    python tests/golden/make_golden_vx3.py

Each fixture holds the reference's state after the scenario's `steps` calls of VX3_VoxelyzeKernel::doTimeStep (scenario
definitions: tests/scenarios.py; the model is regenerated from its seed), floats as C99 hex (exact), integer arrays as
lists.  tests/test_oracle_vs_vx3ref.py::test_oracle_matches_vx3_golden_fixture replays them against the oracle wherever
the library itself is not available (GPU box).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import util  # noqa: E402
from scenarios import SCENARIOS, scenario  # noqa: E402

FLOAT_KEYS = util.KIN + util.LINKF + util.LINKS + ["link_rest_length", "link_strain", "link_max_strain", "link_strain_offset", "link_stress",
                                                   "temp", "contact_force", "signal"]
INT_KEYS = ["vox_flags", "vox_links", "link_vneg", "link_vpos", "link_axis", "link_mat", "link_flags"]

for name in sorted(SCENARIOS):
    sc = scenario(name)
    spec = sc["spec"]()
    lib = util.load_engine()
    b, d = spec.build(lib)
    if sc["link_capacity"]:
        d.contents.link_capacity = sc["link_capacity"]
    ref = util.Vx3RefSim(spec, d)
    dt = float(np.float32(0.9 * ref.recommended_dt())) if sc["dt"] == "fixed" else -1.0
    assert ref.step(sc["steps"], dt) == sc["steps"]
    st = ref.state()
    # keep the files small: every float array of a large scenario is sampled with a fixed stride
    n_float = sum(np.asarray(st[k]).size for k in FLOAT_KEYS)
    out = {"case": name, "steps": sc["steps"], "dt": float(dt).hex(), "covers": sc["covers"],
           "source": "reference VX3 device code (src/VX3/*.cu) compiled for the host, g++ -O2 -ffp-contract=off",
           "state": {k: [float(x).hex() for x in np.asarray(st[k], np.float64).ravel()] for k in FLOAT_KEYS},
           "ints": {k: [int(x) for x in np.asarray(st[k]).ravel()] for k in INT_KEYS}}
    with open(os.path.join(HERE, "vx3_" + name + ".json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(name, ref.counts(), n_float)
    lib.vx3_builder_destroy(b)


# ---- the stdout of the reference's CUDA_Simulation kernel (history frames, real_stepsize line, start / end lines) ----
def history_case(tag, spec, desc, vxa_text=None):
    ref = util.Vx3RefSim(spec, desc, vxa_text=vxa_text)
    out = ref.run_simulation()
    r = ref.result(refresh=False)
    with open(os.path.join(HERE, "vx3_stdout_%s.txt" % tag), "wb") as f:
        f.write(out)
    res = {"steps": int(r.steps), "current_time": float(r.current_time).hex(), "fitness_score": float(r.fitness_score).hex(),
           "current_com": [float(x).hex() for x in r.current_com], "initial_com": [float(x).hex() for x in r.initial_com],
           "num_voxel": int(r.num_voxel), "source": "reference CUDA_Simulation (VX3_SimulationManager.cu:11-121) compiled for the host"}
    with open(os.path.join(HERE, "vx3_stdout_%s.json" % tag), "w") as f:
        json.dump(res, f)
    print(tag, len(out), "bytes of stdout,", r.steps, "steps")


from scenarios import history_spec  # noqa: E402

lib = util.load_engine()
spec = history_spec()
b, d = spec.build(lib)
history_case("runner", spec, d)
lib.vx3_builder_destroy(b)
# BASELINE config 1: the reference's own demos/basic (byte copy under tests/golden/demo_basic/), VXD merged by the product reader
demo = os.path.join(HERE, "demo_basic")
b = lib.vx3_vxa_load(os.path.join(demo, "base.vxa").encode(), os.path.join(demo, "robot.vxd").encode())
assert b, lib.vx3_model_last_error()
d = lib.vx3_builder_build(b)
history_case("demo_basic", None, d, vxa_text=open(os.path.join(demo, "base.vxa")).read())
lib.vx3_builder_destroy(b)
