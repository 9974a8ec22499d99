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
