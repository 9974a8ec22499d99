"""CPU tests of the VXA/VXD front end (csrc/host/vx3_vxa.cpp, vx3_xml.cpp): the same VXA text must give the same
flat model through (a) our reader, (b) the programmatic builder, and (c) the reference's own CVX_Sim reader."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from util import cube_spec, desc_arrays
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200.model import ModelSpec, expr_to_tokens


def parse(lib, vxa, vxd=None, name=b"t"):
    b = lib.vx3_vxa_parse(vxa.encode(), vxd.encode() if vxd else None, name)
    assert b, lib.vx3_model_last_error().decode()
    d = lib.vx3_builder_build(b)
    assert d, lib.vx3_model_last_error().decode()
    return b, d


def assert_same_model(a, b, skip_link_mat_fields=()):
    for k in a:
        if k in ("voxel_mats", "link_mats"):
            assert len(a[k]) == len(b[k]), k
            for i, (x, y) in enumerate(zip(a[k], b[k])):
                for f in x:
                    if k == "link_mats" and f in skip_link_mat_fields:
                        continue
                    same = x[f] == y[f] or (isinstance(x[f], float) and x[f] != x[f] and y[f] != y[f])
                    assert same, "%s[%d].%s: %r != %r" % (k, i, f, x[f], y[f])
        elif isinstance(a[k], np.ndarray):
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        else:
            assert a[k] == b[k], k


def programs(d):
    out = {}
    for s in range(abi.VX3_PROG_COUNT):
        p = d.contents.prog[s]
        out[s] = [(p.tok[i].op, p.tok[i].value) for i in range(p.n)]
    return out


def options(d):
    return {n: getattr(d.contents.opt, n) for n, _ in abi.SimOptions._fields_ if n != "_pad"}


def full_spec():
    spec = cube_spec((4, 3, 3), seed=19, actuated=True, holes=0.2, lift=1, name="full")
    spec.set_options(enable_attach=1, enable_detach=1, safety_guard=77, record_step_size=50, record_link=1,
                     max_dist_in_voxel_lengths_to_count_as_pair=2.5, enable_expansion=1)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.25)))
    spec.set_program(abi.PROG_FITNESS, ("SQRT", ("ADD", ("MUL", ("VAR", "x"), ("VAR", "x")), ("MUL", ("VAR", "y"), ("VAR", "y")))))
    spec.set_program(abi.PROG_FORCE_Z, ("MUL", ("CONST", -0.5), ("SIN", ("VAR", "t"))))
    spec.set_program(abi.PROG_ATTACH_2, ("GREATERTHAN", ("VAR", "z"), ("CONST", 0.01)))
    return spec


def _random(family, seed):
    from scenarios import random_spec, random_spec2
    return lambda: (random_spec if family == "a" else random_spec2)(seed)


@pytest.mark.parametrize("make", [lambda: cube_spec((3, 3, 3), seed=3, actuated=False), lambda: cube_spec((5, 4, 3), seed=5, actuated=True, holes=0.3, lift=1),
                                  full_spec] + [_random("a", k) for k in range(12)] + [_random("b", k) for k in range(12)])
def test_reader_equals_programmatic_builder(make):
    spec = make()
    lib = util.load_engine()
    b1, d1 = spec.build(lib)
    b2, d2 = parse(lib, spec.to_vxa(), name=spec.name.encode())
    try:
        assert_same_model(desc_arrays(d1), desc_arrays(d2))
        assert programs(d1) == programs(d2)
        assert options(d1) == options(d2)
        for s, expr in spec.programs.items():
            assert programs(d2)[s] == [(op, v) for op, v in expr_to_tokens(expr)]
    finally:
        lib.vx3_builder_destroy(b1)
        lib.vx3_builder_destroy(b2)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built (no reference tree on this box)")
def test_reader_equals_reference_reader_on_shipped_demo():
    """demos/basic/base.vxa: our reader vs CVX_Sim::LoadVXAFile + Import (no <Environment>, no <Damping>: defaults)."""
    path = os.path.join(util.REFERENCE_TREE, "demos", "basic", "base.vxa")
    if not os.path.exists(path):
        pytest.skip("reference demo not present")
    lib = util.load_engine()
    ref = util.load_ref()
    h = ref.vxref_load_vxa(path.encode())
    assert ref.vxref_ok(h)
    dref = ref.vxref_export(h)
    b = lib.vx3_vxa_load(path.encode(), None)
    assert b, lib.vx3_model_last_error().decode()
    d = lib.vx3_builder_build(b)
    try:
        assert_same_model(desc_arrays(d), desc_arrays(dref), skip_link_mat_fields=("matid",))
        # the demo's stop condition t - 1 > 0: tokens [CONST 1, VAR t, SUB, END]
        assert programs(d)[abi.PROG_STOP] == [(abi.OP["CONST"], 1.0), (abi.OP["VAR"], 4.0), (abi.OP["SUB"], 0.0), (abi.OP["END"], 0.0)]
        assert d.contents.opt.record_step_size == 100 and d.contents.opt.enable_collision == 1
        assert d.contents.name == b"base.vxa"
    finally:
        lib.vx3_builder_destroy(b)


def test_vxd_replace_merges_subtrees():
    spec = cube_spec((2, 2, 2), seed=1, actuated=False, name="base")
    other = cube_spec((3, 2, 1), seed=2, actuated=False, name="other")
    vxa = spec.to_vxa()
    o = other.to_vxa()
    structure = o[o.index("<Structure"):o.index("</Structure>") + len("</Structure>")].replace("<Structure ", '<Structure replace="VXA.VXC.Structure" ', 1)
    vxd = "<VXD>\n%s\n<DtFrac replace=\"VXA.Simulator.Integration.DtFrac\">0.5</DtFrac>\n<RawPrint>ignored: no replace attribute</RawPrint>\n</VXD>" % structure
    lib = util.load_engine()
    b, d = parse(lib, vxa, vxd, b"robot_7.vxd")
    bo, do = other.build(lib)
    try:
        a, c = desc_arrays(d), desc_arrays(do)
        for k in ("ix", "iy", "iz", "vox_mat", "link_vneg", "link_vpos", "link_axis"):
            np.testing.assert_array_equal(a[k], c[k], err_msg=k)
        assert d.contents.opt.dt_frac == 0.5
        assert d.contents.name == b"robot_7.vxd"
    finally:
        lib.vx3_builder_destroy(b)
        lib.vx3_builder_destroy(bo)


def test_malformed_inputs_fail_loudly():
    lib = util.load_engine()
    for bad in ("<VXA><Simulator></VXA>", "<VXA></VXA>", "not xml", "<VXA><VXC><Structure Compression=\"ZLIB\"/></VXC></VXA>"):
        assert not lib.vx3_vxa_parse(bad.encode(), None, None)
        assert lib.vx3_model_last_error()
    spec = cube_spec((2, 2, 2))
    broken = spec.to_vxa().replace("<mtSUB>", "<mtBOGUS>").replace("</mtSUB>", "</mtBOGUS>")
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 1)))
    broken = spec.to_vxa().replace("mtSUB", "mtBOGUS")
    assert not lib.vx3_vxa_parse(broken.encode(), None, None)
    assert b"not implemented" in lib.vx3_model_last_error()


def test_report_writer(tmp_path):
    lib = C.CDLL(util.load_engine()._name)
    lib.vx3_write_report.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(abi.Result), C.c_int]
    arr = (abi.Result * 2)()
    arr[0].name, arr[0].fitness_score, arr[0].num_voxel, arr[0].current_time = b"robot_3.vxd", 1.5, 12, 0.25
    arr[1].name, arr[1].fitness_score = b"robot_1.vxd", float("nan")
    out = tmp_path / "r.vxr"
    assert lib.vx3_write_report(str(out).encode(), b"/some/where/gen_12", arr, 2) == 0
    text = out.read_text()
    assert "<inputdir>gen_12</inputdir>" in text
    assert "<bestfit><filename>robot_3.vxd</filename><fitness_score>1.5</fitness_score></bestfit>" in text
    assert "<robot_3><currentTime>0.25</currentTime><fitness_score>1.5</fitness_score><num_voxel>12</num_voxel>" in text
    assert "<robot_1>" in text and "nan" in text
