#!/usr/bin/env python
"""CPU-only sweep: seeded random models (tests/scenarios.random_spec) through the oracle AND the reference's own VX3 code compiled
for the host (oracle/_ref/libvxref_vx3.so), bit for bit.  Run where /root/reference exists:

    python scripts/fuzz_oracle_vs_reference.py 8 60        # seeds [8, 60) of random_spec
    python scripts/fuzz_oracle_vs_reference.py 0 40 b      # ... of random_spec2 (signals, cilia, removal, programs, targets)

A seed that differs is a finding: pin it as a named scenario in tests/scenarios.py, fix the oracle, then the engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import util  # noqa: E402
from scenarios import random_spec, random_spec2, random_spec3  # noqa: E402
from test_oracle_vs_vx3ref import STATE_KEYS  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
family = sys.argv[3] if len(sys.argv) > 3 else "a"  # a: random_spec, b: random_spec2 (optional physics), c: random_spec3 (everything at once)
bad = []
for seed in range(lo, hi):
    spec = {"a": random_spec, "b": random_spec2, "c": random_spec3}[family](seed)
    lib = util.load_engine()
    b, d = spec.build(lib)
    d.contents.link_capacity = d.contents.n_links + 4096
    try:
        ref, orc = util.Vx3RefSim(spec, d), util.OracleSim(d)
        dt = -1.0 if (seed % 2 or family == "c") else float(np.float32(0.9 * orc.recommended_dt()))
        steps, chunk = (1600, 400) if (seed % 2 or family != "a") else (600, 150)
        done, ok = 0, True
        while done < steps and ok:
            ref.step(chunk, dt)
            orc.step(chunk, dt)
            done += chunk
            try:
                util.assert_bit_equal(orc.state(), ref.state(), STATE_KEYS, "seed %d after %d steps" % (seed, done))
            except AssertionError as e:
                ok = False
                bad.append(seed)
                print("MISMATCH", str(e)[:300])
        print("seed", seed, "ok" if ok else "BAD", orc.counts())
    finally:
        lib.vx3_builder_destroy(b)
print("mismatching seeds:", bad)
