#!/usr/bin/env python
"""developer helper: wall-clock breakdown of the e2e path (create / step / results / positions / destroy)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
graft.load_package()
from voxcraft_sim_b200 import workloads as W
from voxcraft_sim_b200.engine import Batch
from voxcraft_sim_b200.libs import load_engine
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
lib = load_engine(False)
specs = [W.c2_spec()] if wl == "c2" else [W.c3_spec(k) for k in range(512)]
built = [s.build(lib) for s in specs]
descs = [d for _, d in built]
for it in range(4):
    if it == 3:
        os.environ["VX3_CREATE_TIMING"] = "1"
    t = [time.perf_counter()]
    bt = Batch(descs); t.append(time.perf_counter())
    bt.step(1000); t.append(time.perf_counter())
    bt.results(); t.append(time.perf_counter())
    for i in range(len(descs)):
        bt.positions(i)
    t.append(time.perf_counter())
    bt.close(); t.append(time.perf_counter())
    print(wl, "iter", it, " ".join("%s %.2f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(("create", "step", "results", "positions", "destroy"), t, t[1:])), "device %.2f ms" % 0.0, flush=True)
