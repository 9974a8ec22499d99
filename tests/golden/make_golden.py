"""Generates tests/golden/*.json from the UNMODIFIED reference CPU library (oracle/_ref/libvxref.so, compiled in
place from /root/reference by oracle/Makefile).  Run where the reference tree exists:

    python tests/golden/make_golden.py

Each fixture holds the reference's state after N calls of CVoxelyze::doTimeStep(dt) on a seeded model (the model
itself is regenerated from the seed by tests/test_oracle_vs_ref.py::make_spec), floats as C99 hex (exact).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import util  # noqa: E402
import test_oracle_vs_ref as T  # noqa: E402

STEPS = 600

for name in sorted(T.CASES):
    spec = T.make_spec(name)
    ref = util.RefSim(spec)
    dt = float(np.float32(0.9 * ref.recommended_dt()))
    assert ref.step(STEPS, dt) == STEPS
    st = ref.state()
    out = {"case": name, "steps": STEPS, "dt": dt.hex(), "n_voxels": ref.nv, "n_links": ref.nl,
           "source": "reference CPU library src/old (CVoxelyze::doTimeStep), g++ -O2 -ffp-contract=off",
           "state": {k: [float(x).hex() for x in np.asarray(st[k], np.float64).ravel()] for k in util.KIN + util.LINKF + util.LINKS}}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f)
    print(name, ref.nv, ref.nl)
