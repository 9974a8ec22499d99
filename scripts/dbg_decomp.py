import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import __graft_entry__ as graft
graft.load_package()
import util
from util import cube_spec, EngineBatch
from voxcraft_sim_b200 import parallel
world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec = cube_spec((9, 4, 3), seed=21, actuated=True, holes=0.1, name="decomp")
lib = util.load_engine()
b, d = spec.build(lib)
dt = float(np.float32(0.9 * lib.vx3_model_recommended_dt(d)))
whole = EngineBatch([d]); whole.set_profiling(False, use_persistent=False)
slabs = [parallel.partition_slabs(d, world, r) for r in range(world)]
parts = [parallel.DecomposedBody(s, dt) for s in slabs]
for r in range(world):
    for side, nb in ((0, r - 1), (1, r + 1)):
        if 0 <= nb < world:
            parts[r].batch.halo_connect_local(side, parts[nb].batch)
ix = np.ctypeslib.as_array(d.contents.ix, shape=(d.contents.n_voxels,))
for s in slabs:
    print("rank", s.rank, "nvox", len(s.voxels), "owned", int(s.owned.sum()), "x owned", sorted(set(ix[s.voxels[s.owned]].tolist())), "x ghost", sorted(set(ix[s.voxels[~s.owned]].tolist())),
          "send", {k: len(v) for k, v in s.send.items()}, "recv", {k: len(v) for k, v in s.recv.items()})
for step in range(1, 8):
    whole.step(1, dt)
    for p in parts: p.batch.step_async(1, dt)
    for p in parts: p.batch.sync()
    sw = whole.state(0)
    for s, p in zip(slabs, parts):
        sp = p.batch.state(0)
        for key, w in (("pos", 3), ("orient", 4), ("lin_mom", 3)):
            a = np.asarray(sp[key]).reshape(-1, w)
            ref = np.asarray(sw[key]).reshape(-1, w)[s.voxels]
            bad = np.nonzero((a != ref).any(axis=1))[0]
            if len(bad):
                print("step", step, "rank", s.rank, key, "mismatch at local", bad[:10].tolist(), "owned?", s.owned[bad[:10]].tolist(), "x", ix[s.voxels[bad[:10]]].tolist(),
                      "maxabs", float(np.abs(a - ref).max()))
