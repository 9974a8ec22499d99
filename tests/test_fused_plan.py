"""Host-side block partition of the fused step (vx3_fused.cuh, fused_plan_build) through the C ABI's CPU hook: every voxel
in exactly one block, every link either interior to exactly one block (both ends in it, at the local positions its entry
names) or exactly once on the face list, blocks within the size bound and inside one simulation."""
import ctypes as C

import pytest

import util
from util import cube_spec
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200 import workloads as W


def plan_check(lib, descs, bv):
    arr = (abi.ModelDesc * len(descs))()
    for i, d in enumerate(descs):
        C.memmove(C.byref(arr[i]), d, C.sizeof(abi.ModelDesc))
    out = (C.c_int64 * 8)()
    rc = lib.vx3_fused_plan_check(arr, len(descs), bv, out)
    assert rc == 0, lib.vx3_last_error().decode()
    return list(out)


@pytest.mark.parametrize("bv", [8, 30, 0])
def test_partition_covers_every_voxel_and_link_once(bv):
    lib = util.load_engine()
    specs = [cube_spec((7, 6, 5), seed=17, holes=0.2, name="p0"), cube_spec((1, 1, 1), seed=1, name="p1"), cube_spec((12, 3, 2), seed=5, name="p2"),
             W.c3_spec(3), W.c3_spec(11)]
    built = [s.build(lib) for s in specs]
    try:
        descs = [d for _, d in built]
        blocks, interior, face, links, voxels, largest, vok, lok = plan_check(lib, descs, bv)
        assert voxels == sum(d.contents.n_voxels for d in descs) and links == sum(d.contents.n_links for d in descs)
        assert vok == voxels, "every voxel in exactly one block (negative: inconsistent entries)"
        assert lok == links and interior + face == links
        assert largest <= (bv if bv > 0 else 120)
        assert blocks >= len(descs)
        if bv == 0:
            assert interior > 0.6 * links  # compact blocks: most links are interior
    finally:
        for b, _ in built:
            lib.vx3_builder_destroy(b)


def test_large_lattice_blocks_are_compact():
    lib = util.load_engine()
    b, d = W.c5_spec((40, 40, 30)).build(lib)
    try:
        blocks, interior, face, links, voxels, largest, vok, lok = plan_check(lib, [d], 0)
        assert vok == voxels == 48000 and lok == links
        assert largest <= 120 and blocks <= voxels // 100
        assert face < 0.3 * links, "face links %d of %d" % (face, links)
    finally:
        lib.vx3_builder_destroy(b)
