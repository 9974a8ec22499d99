"""GPU: the stdout stream of a run — .history header, `real_stepsize:` line, voxel / link frames, start / end lines — byte for
byte against what the reference's OWN CUDA_Simulation kernel printed for the same input (fixtures tests/golden/vx3_stdout_*.txt,
captured from src/VX3/VX3_SimulationManager.cu:11-121 compiled for the host by oracle/ref_vx3; generator
tests/golden/make_golden_vx3.py), through vx3_batch_run's history callback and through the voxcraft-sim executable."""
import json
import os
import subprocess

import numpy as np
import pytest

import util
from scenarios import history_spec
from util import EngineBatch

pytestmark = pytest.mark.gpu

LIBDIR = os.path.join(util.ROOT, "voxcraft-sim_b200", "lib")
DEMO = os.path.join(util.GOLDEN, "demo_basic")


def golden(tag):
    out = open(os.path.join(util.GOLDEN, "vx3_stdout_%s.txt" % tag), "rb").read()
    res = json.load(open(os.path.join(util.GOLDEN, "vx3_stdout_%s.json" % tag)))
    return out, res


def first_difference(a, b):
    n = min(len(a), len(b))
    i = next((k for k in range(n) if a[k] != b[k]), n)
    return "byte %d of %d/%d: ...%r vs ...%r" % (i, len(a), len(b), a[max(0, i - 60):i + 40], b[max(0, i - 60):i + 40])


def check_result(r, res):
    assert r.steps == res["steps"]
    assert r.current_time == float.fromhex(res["current_time"])
    assert r.num_voxel == res["num_voxel"]
    np.testing.assert_allclose(r.fitness_score, float.fromhex(res["fitness_score"]), rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(list(r.current_com), [float.fromhex(x) for x in res["current_com"]], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(list(r.initial_com), [float.fromhex(x) for x in res["initial_com"]], rtol=1e-12, atol=1e-18)


def run_and_collect(desc):
    eng = EngineBatch([desc])
    chunks = []
    eng.run(history=chunks)
    r = eng.results()[0]
    eng.close()
    assert all(sim == 0 for sim, _ in chunks)
    return b"".join(c for _, c in chunks), r


def test_history_stream_equals_the_reference_stdout():
    want, res = golden("runner")
    spec = history_spec()
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        got, r = run_and_collect(d)
        assert got == want, first_difference(got, want)
        assert got.count(b"<<<Step") >= 3 and got.count(b"|[[[") >= 3 and b"real_stepsize: " in got
        check_result(r, res)
    finally:
        lib.vx3_builder_destroy(b)


def test_demo_basic_config1_through_the_engine():
    """BASELINE config 1: the reference's demos/basic VXA + VXD (byte copies under tests/golden/demo_basic) through the
    product's VXA/VXD reader and vx3_batch_run: 6,982 steps, identical stdout, identical end state."""
    want, res = golden("demo_basic")
    lib = util.load_engine()
    b = lib.vx3_vxa_load(os.path.join(DEMO, "base.vxa").encode(), os.path.join(DEMO, "robot.vxd").encode())
    assert b, lib.vx3_model_last_error()
    d = lib.vx3_builder_build(b)
    try:
        got, r = run_and_collect(d)
        assert got == want, first_difference(got, want)
        check_result(r, res)
        assert r.steps == 6982
    finally:
        lib.vx3_builder_destroy(b)


def test_demo_basic_config1_through_voxcraft_sim(tmp_path):
    """The same demo through the drop-in executables: `voxcraft-sim -i demos/basic -o report.xml` (stdout redirected = the
    .history file voxcraft-viz reads).  Every line the reference prints must be in the stream, in order."""
    want, res = golden("demo_basic")
    report = tmp_path / "report.xml"
    p = subprocess.run([os.path.join(LIBDIR, "voxcraft-sim"), "-i", DEMO, "-o", str(report), "-w", os.path.join(LIBDIR, "vx3_node_worker"), "-f"],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout
    pos = 0
    for line in want.split(b"\n"):
        if not line:
            continue
        k = p.stdout.find(line, pos)
        assert k >= 0, "missing or out of order in the executable's stdout: %r" % line[:120]
        pos = k + len(line)
    text = report.read_text()
    assert "<bestfit><filename>robot.vxd</filename>" in text
    import re
    m = re.search(r"<robot><currentTime>([^<]+)</currentTime><fitness_score>([^<]+)</fitness_score><num_voxel>1</num_voxel>", text)
    assert m, text
    assert float(m.group(1)) == float.fromhex(res["current_time"])


def test_report_carries_per_voxel_positions(tmp_path):
    """SavePositionOfAllVoxels: <init_pos> / <pos> / <mats> of vx3_node_worker.cu:122-139 (std::to_string format)."""
    import re
    from test_gpu_worker import robot
    spec = robot(100)
    spec.set_options(save_position_of_all_voxels=1, record_step_size=0)
    gen = tmp_path / "gen"
    gen.mkdir()
    (gen / "base.vxa").write_text(spec.to_vxa())
    (gen / "a.vxd").write_text("<VXD>\n</VXD>\n")
    report = tmp_path / "r.xml"
    p = subprocess.run([os.path.join(LIBDIR, "voxcraft-sim"), "-i", str(gen), "-o", str(report), "-w", os.path.join(LIBDIR, "vx3_node_worker"), "-f"],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout
    text = report.read_text()
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc = util.OracleSim(d)
        orc.run()
        so = orc.state()
        nv = d.contents.n_voxels
        m = re.search(r"<init_pos>([^<]*)</init_pos><pos>([^<]*)</pos><mats>([^<]*)</mats>", text)
        assert m, text[:2000]
        init = np.array([[float(x) for x in t.split(",")] for t in m.group(1).strip(";").split(";")])
        pos = np.array([[float(x) for x in t.split(",")] for t in m.group(2).strip(";").split(";")])
        mats = [int(x) for x in m.group(3).strip(";").split(";")]
        assert init.shape == pos.shape == (nv, 3) and len(mats) == nv
        np.testing.assert_allclose(pos, so["pos"], atol=1.01e-6)   # "%f": six decimals
        want_init = np.ctypeslib.as_array(d.contents.pos, shape=(nv * 3,)).reshape(nv, 3)
        np.testing.assert_allclose(init, want_init, atol=1.01e-6)
        assert mats == [d.contents.voxel_mats[d.contents.vox_mat[i]].matid for i in range(nv)]
        assert re.fullmatch(r"(-?\d+\.\d{6},-?\d+\.\d{6},-?\d+\.\d{6};)+", m.group(2))
    finally:
        lib.vx3_builder_destroy(b)
