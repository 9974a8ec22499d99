#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (scripts/gpu.sh sanitize):
persistent on-chip kernel, fused block step, streaming link / voxel passes, collision grid + contact + attach / detach
resolution, signals, SecondaryExperiment, and the halo exchange between two slabs driven by one process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

graft.load_package()
import util  # noqa: E402
from scenarios import scenario  # noqa: E402
from util import EngineBatch  # noqa: E402
from voxcraft_sim_b200 import parallel  # noqa: E402


def run(name, steps, persistent, fused=False):
    sc = scenario(name)
    spec = sc["spec"]()
    lib = util.load_engine()
    b, d = spec.build(lib)
    if sc["link_capacity"]:
        d.contents.link_capacity = sc["link_capacity"]
    if fused:
        os.environ["VX3_FUSED"] = "1"
    eng = EngineBatch([d, d])
    os.environ.pop("VX3_FUSED", None)
    eng.set_profiling(False, use_persistent=persistent)
    eng.step(steps)
    r = eng.results()[0]
    print("sanitize case %-16s %s steps %d links %d" % (name, "persistent" if persistent else ("fused" if fused else "streaming"), r.steps, r.num_links), flush=True)
    eng.close()
    lib.vx3_builder_destroy(b)


def halo_case(steps=60):
    spec = util.cube_spec((9, 4, 3), seed=21, actuated=True, holes=0.1, name="halo")
    lib = util.load_engine()
    b, d = spec.build(lib)
    dt = float(np.float32(0.9 * lib.vx3_model_recommended_dt(d)))
    slabs = [parallel.partition_slabs(d, 2, r) for r in range(2)]
    parts = [parallel.DecomposedBody(s, dt) for s in slabs]
    parts[0].batch.halo_connect_local(1, parts[1].batch)
    parts[1].batch.halo_connect_local(0, parts[0].batch)
    for _ in range(steps // 20):
        for p in parts:
            p.batch.step_async(20, dt)
        for p in parts:
            p.batch.sync()
    print("sanitize case halo (2 slabs, one process) steps", steps, flush=True)
    for p in parts:
        p.batch.close()
    lib.vx3_builder_destroy(b)


if __name__ == "__main__":
    single = scenario("act333")["spec"]()
    lib = util.load_engine()
    b, d = single.build(lib)
    eng = EngineBatch([d])  # one body: the persistent on-chip kernel
    eng.step(120)
    print("sanitize case act333 persistent steps", eng.results()[0].steps, flush=True)
    eng.close()
    lib.vx3_builder_destroy(b)
    run("ragged", 100, persistent=False, fused=True)
    run("act333", 100, persistent=False)
    run("pile_sticky", 2500, persistent=False)
    run("detach", 1500, persistent=False)
    run("sig_body", 100, persistent=False)
    run("secondary", 200, persistent=False)
    # the halo exchange is NOT run here: compute-sanitizer serialises kernels, and a receive kernel that spins for a send
    # kernel queued behind it on the same device never returns (halo_case needs two devices; tests/test_decomposition.py and
    # scripts/check_decomp_mp.py cover it without the tool)
    if "--halo" in sys.argv:
        halo_case()
    print("SANITIZE CASES DONE")
