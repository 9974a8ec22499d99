"""GPU test of the drop-in executables: voxcraft-sim -> vx3_node_worker -> .vxr report + .history stream, checked
against the oracle running the same VXA+VXD inputs (fitness = CoM displacement, within 1e-6 relative)."""
import os
import re
import subprocess

import numpy as np
import pytest

import util
from util import OracleSim, cube_spec
from voxcraft_sim_b200 import abi

pytestmark = pytest.mark.gpu

LIBDIR = os.path.join(util.ROOT, "voxcraft-sim_b200", "lib")


def robot(seed):
    spec = cube_spec((3, 3, 2), seed=seed, actuated=True, holes=0.15, name="robot_%d" % seed)
    spec.set_env(temp_period=0.02)
    spec.set_options(record_step_size=100, record_link=1)
    spec.set_program(abi.PROG_STOP, ("SUB", ("VAR", "t"), ("CONST", 0.06)))
    spec.set_program(abi.PROG_FITNESS, ("SQRT", ("ADD", ("MUL", ("VAR", "x"), ("VAR", "x")), ("MUL", ("VAR", "y"), ("VAR", "y")))))
    return spec


def test_voxcraft_sim_end_to_end(tmp_path):
    base = robot(100)
    gen = tmp_path / "gen_0"
    gen.mkdir()
    (gen / "base.vxa").write_text(base.to_vxa())
    seeds = [101, 102, 103]
    for s in seeds:
        o = robot(s).to_vxa()
        structure = o[o.index("<Structure"):o.index("</Structure>") + len("</Structure>")].replace("<Structure ", '<Structure replace="VXA.VXC.Structure" ', 1)
        (gen / ("robot_%d.vxd" % s)).write_text("<VXD>\n%s\n</VXD>\n" % structure)
    report = tmp_path / "report.xml"
    p = subprocess.run([os.path.join(LIBDIR, "voxcraft-sim"), "-i", str(gen), "-o", str(report), "-w", os.path.join(LIBDIR, "vx3_node_worker"), "-f"],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout
    assert report.exists(), p.stdout
    text = report.read_text()
    assert "<inputdir>gen_0</inputdir>" in text
    # oracle on the same merged inputs (base material palette + the VXD's structure)
    lib = util.load_engine()
    fits = {}
    for s in seeds:
        vxd = (gen / ("robot_%d.vxd" % s)).read_text()
        b = lib.vx3_vxa_parse(base.to_vxa().encode(), vxd.encode(), ("robot_%d.vxd" % s).encode())
        d = lib.vx3_builder_build(b)
        orc = OracleSim(d)
        orc.run()
        r = orc.result(refresh=False)
        fits[s] = (r.fitness_score, r.current_time, r.num_voxel)
        lib.vx3_builder_destroy(b)
    for s in seeds:
        m = re.search(r"<robot_%d><currentTime>([^<]+)</currentTime><fitness_score>([^<]+)</fitness_score><num_voxel>(\d+)</num_voxel>" % s, text)
        assert m, text
        t, fit, nv = float(m.group(1)), float(m.group(2)), int(m.group(3))
        assert nv == fits[s][2]
        assert t == fits[s][1]
        np.testing.assert_allclose(fit, fits[s][0], rtol=1e-6)
    best = max(seeds, key=lambda s: fits[s][0])
    assert "<bestfit><filename>robot_%d.vxd</filename>" % best in text
    # detail entries are sorted by fitness, descending (sortResults)
    order = [int(x) for x in re.findall(r"<robot_(\d+)><currentTime>", text)]
    assert order == sorted(seeds, key=lambda s: -fits[s][0])
    # history stream on stdout: header + frames (VX3_SimulationManager.cu:40-50,70-114)
    assert "{{{setting}}}<rescale>0.001</rescale>" in p.stdout
    assert re.search(r"<<<Step0 Time:[0-9.]+>>>[-0-9.,;]+<<<>>>\|\[\[\[0\]\]\][-0-9.,;]+\[\[\[\]\]\]", p.stdout)
    assert p.stdout.count("<<<Step") >= 3 * 5
