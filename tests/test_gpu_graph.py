"""GPU: the CUDA-Graph stretches of the streaming path (vx3_engine.cu, advance): 32 plain steps captured once and replayed
must be bit-identical to launching every step's kernels one by one (VX3_GRAPH=0), on batches that exercise every kernel a
step can contain — links, voxels, collision grid, contact, attach / detach, signals, SecondaryExperiment — and across
centre-of-mass sampling steps (which never enter a graph)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, json, hashlib
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np
import util
from scenarios import scenario
from util import EngineBatch
out = {}
for name in %(names)r:
    sc = scenario(name)
    spec = sc["spec"]()
    lib = util.load_engine()
    b, d = spec.build(lib)
    if sc["link_capacity"]:
        d.contents.link_capacity = sc["link_capacity"]
    eng = EngineBatch([d, d])   # two simulations per batch
    eng.set_profiling(False, use_persistent=False)
    eng.step(sc["steps"])
    h = hashlib.sha256()
    for sim in (0, 1):
        st = eng.state(sim, link_cap=sc["link_capacity"] or None)
        for k in sorted(st):
            h.update(np.ascontiguousarray(st[k]).tobytes())
    r = eng.results()[0]
    out[name] = [h.hexdigest(), int(r.steps), int(r.num_links), eng.timing()[1]]
    eng.close()
print("RESULT " + json.dumps(out))
"""

NAMES = ["act333", "ragged", "pile_sticky", "detach", "secondary", "sig_body", "closeness", "poisson_bilinear"]


def run_child(graph):
    env = dict(os.environ, VX3_GRAPH="1" if graph else "0")
    code = CHILD % dict(root=util.ROOT, tests=os.path.join(util.ROOT, "tests"), names=NAMES)
    p = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout
    import json
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_graph_replay_is_bit_identical_to_per_step_launches():
    with_graph, without = run_child(True), run_child(False)
    for name in NAMES:
        assert with_graph[name][:3] == without[name][:3], name
        assert with_graph[name][3] == without[name][3], "%s: kernel launch count differs" % name
