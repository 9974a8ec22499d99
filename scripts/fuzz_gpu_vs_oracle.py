#!/usr/bin/env python
"""GPU sweep: seeded random models (tests/scenarios.random_spec) through the engine (C ABI) and the CPU oracle; topology bit-exact,
kinematic state to 1e-8 of the array's scale (a screening bar: a seed that fails goes through the full gate of
tests/test_gpu_scenarios.py as a named scenario).

    python scripts/fuzz_gpu_vs_oracle.py 8 48
    python scripts/fuzz_gpu_vs_oracle.py 0 40 b     # random_spec2: signals (state bit-exact), cilia, removal, programs, targets"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import util  # noqa: E402
from scenarios import random_spec, random_spec2, random_spec3  # noqa: E402

INT_KEYS = ["vox_flags", "vox_links", "link_vneg", "link_vpos", "link_axis", "link_flags"]
lo, hi = int(sys.argv[1]), int(sys.argv[2])
family = sys.argv[3] if len(sys.argv) > 3 else "a"  # a: random_spec, b: random_spec2 (optional physics), c: random_spec3 (everything at once)
bad = []
for seed in range(lo, hi):
    spec = {"a": random_spec, "b": random_spec2, "c": random_spec3}[family](seed)
    lib = util.load_engine()
    b, d = spec.build(lib)
    cap = d.contents.n_links + 4096
    d.contents.link_capacity = cap
    try:
        eng, orc = util.EngineBatch([d]), util.OracleSim(d)
        dt = -1.0 if (seed % 2 or family == "c") else float(np.float32(0.9 * orc.recommended_dt()))
        steps, chunk = (1600, 400) if (seed % 2 or family != "a") else (600, 150)
        done, why = 0, None
        while done < steps and why is None:
            eng.step(chunk, dt) if dt > 0 else eng.step(chunk)
            orc.step(chunk, dt)
            done += chunk
            se, so = eng.state(0, link_cap=cap), orc.state()
            if se["link_vneg"].shape != so["link_vneg"].shape:
                why = "link count %s vs %s" % (se["link_vneg"].shape, so["link_vneg"].shape)
                break
            for k in INT_KEYS + ["signal"]:
                if not np.array_equal(se[k], so[k]):
                    why = "%s after %d steps" % (k, done)
                    break
            for k in ["pos", "orient", "lin_mom"]:
                a, r = np.asarray(se[k], float), np.asarray(so[k], float)
                e = np.abs(a - r).max() / max(np.abs(r).max(), 1e-300)
                if why is None and e > 1e-8:
                    why = "%s rel err %.2e after %d steps" % (k, e, done)
            re, ro = eng.results()[0], orc.result()
            if why is None and abs(re.current_time - ro.current_time) > 1e-12 * abs(ro.current_time):
                why = "current_time %r vs %r" % (re.current_time, ro.current_time)
        print("seed", seed, "ok" if why is None else "BAD " + why, orc.counts(), flush=True)
        if why:
            bad.append(seed)
        eng.close()
    finally:
        lib.vx3_builder_destroy(b)
print("failing seeds:", bad)
