"""Pins the CPU oracle (oracle/vx3_oracle.cpp) and the host model builder against the UNMODIFIED reference CPU
library (src/old + src/VXA compiled in place into oracle/_ref/libvxref.so):

  * the model builder's flat vx3_model_desc == the one exported from the reference's own CVX_Sim::Import of the
    same VXA text (voxel order, link order, link-material table, every derived constant), bit for bit;
  * the oracle in cpu_lib_mode reproduces CVoxelyze::doTimeStep (src/old/Voxelyze.cpp:251-284) bit for bit on
    the feature subset the two share (link forces, integration, floor, friction, static temperature).

When oracle/_ref is absent (the GPU box has no /root/reference) the same checks run against the committed
fixtures in tests/golden/, generated here by tests/golden/make_golden.py from the reference library.
"""
import json
import os

import numpy as np
import pytest

import util
from util import KIN, LINKF, LINKS, OracleSim, RefSim, assert_bit_equal, cube_spec, desc_arrays
from voxcraft_sim_b200.model import ModelSpec

CASES = {
    "cube3": dict(n=(3, 3, 3), seed=3, actuated=False),
    "ragged": dict(n=(5, 4, 3), seed=5, actuated=False, holes=0.3, lift=1),
    "tower": dict(n=(2, 2, 6), seed=9, actuated=False),
    "static_temp": dict(n=(3, 2, 2), seed=13, actuated=False, static_temp=7.5),
    # MatModel 2 / 3 next to a linear material: every pair goes through the springs-in-series blend (VX_MaterialLink.cpp:72-104);
    # soft enough that the body's own weight + the drop drive links past the yield points (loading, unloading, reloading)
    "nonlinear_mix": dict(n=(3, 3, 4), seed=17, actuated=False, lift=1, nonlinear=True),
}
# builder-only cases (the CPU library re-evaluates the transverse strains of a nu != 0 material every step, VX_Link.cpp:154;
# VX3 freezes them, VX3_Link.cu:147-150, and so does the oracle — stepping with nu != 0 is pinned on the VX3 code instead,
# tests/test_oracle_vs_vx3ref.py::poisson_*)
BUILD_ONLY = {
    "poisson_mix": dict(n=(3, 3, 3), seed=19, actuated=False, poisson=True),
    "poisson_nonlinear": dict(n=(3, 3, 3), seed=23, actuated=False, poisson=True, nonlinear=True),
    "poisson_off": dict(n=(2, 2, 2), seed=29, actuated=False, poisson=True, volume_effects=False),
}


# members the reference never initialises for link materials (CVX_MaterialLink's ctor leaves the VX3 additions of
# CVX_Material unset; they are read from zeroed/garbage storage) and the step loop never reads from a link material
LINKMAT_UNSET = {"matid"}


def make_spec(name):
    kw = dict(CASES[name] if name in CASES else BUILD_ONLY[name])
    st = kw.pop("static_temp", None)
    nonlinear, poisson, volume = kw.pop("nonlinear", False), kw.pop("poisson", False), kw.pop("volume_effects", True)
    spec = cube_spec(name=name, **kw)
    if nonlinear:
        spec.materials[1].update(mat_model=2, elastic_mod=4e4, plastic_mod=8e3, yield_stress=12.0, fail_stress=400.0)
        spec.materials[2].update(mat_model=3, elastic_mod=0.0, strain_data=[0.0, 0.0004, 0.01, 0.05, 0.3], stress_data=[0.0, 12.0, 150.0, 300.0, 500.0])
        spec.materials[0].update(elastic_mod=3e4)
    if poisson:
        for m, nu in zip(spec.materials, (0.3, 0.2, 0.0)):
            m["poissons_ratio"] = nu
        spec.set_env(volume_effects_enabled=int(volume))
    if st is not None:  # constant (non-varying) temperature: applied once at Import by both libraries
        spec.set_env(temp_enabled=1, vary_temp_enabled=0, temp_amplitude=st, temp_base=25.0)
    return spec


def demo_spec():
    """demos/basic/base.vxa restated: one passive voxel dropped from z=4 voxel lengths... (1 voxel, 0 links)."""
    spec = ModelSpec(0.01, "demo")
    spec.add_material(elastic_mod=1e4, density=1e3, u_static=1.0, u_dynamic=0.5)
    spec.set_structure(np.ones((1, 1, 1), np.uint8))
    return spec


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built (no reference tree on this box)")
@pytest.mark.parametrize("name", sorted(CASES) + sorted(BUILD_ONLY) + ["random_a%d" % k for k in range(16)] + ["random_b%d" % k for k in range(8)])
def test_builder_equals_reference_import(name):
    """The host builder's model (what the engine consumes) against the reference's CVX_Sim::LoadVXAFile + Import of the same VXA
    text, every array and every material constant bit for bit — on the hand-made cases and on seeded random models
    (tests/scenarios.random_spec / random_spec2: random palettes incl. linear-with-failure and bilinear materials, nu, CTE, holes,
    externals, two-body structures)."""
    if name.startswith("random_"):
        from scenarios import random_spec, random_spec2
        spec = (random_spec if name[7] == "a" else random_spec2)(int(name[8:]))
    else:
        spec = make_spec(name)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        ref = RefSim(spec)
        mine, theirs = desc_arrays(d), desc_arrays(ref.desc)
        for k in mine:
            if k in ("voxel_mats", "link_mats"):
                assert len(mine[k]) == len(theirs[k]), k
                for i, (a, c) in enumerate(zip(mine[k], theirs[k])):
                    for f in a:
                        if k == "link_mats" and f in LINKMAT_UNSET:
                            continue
                        va, vc = a[f], c[f]
                        same = va == vc or (isinstance(va, float) and va != va and vc != vc)
                        assert same, "%s[%d].%s: %r != %r" % (k, i, f, va, vc)
            elif isinstance(mine[k], np.ndarray):
                np.testing.assert_array_equal(mine[k], theirs[k], err_msg=k)
            else:
                assert mine[k] == theirs[k], k
        assert lib.vx3_model_recommended_dt(d) > 0
    finally:
        lib.vx3_builder_destroy(b)


def passive_random_spec(seed):
    """tests/scenarios.random_spec within what the CPU library can do: no per-voxel phase actuation, nu = 0."""
    from scenarios import random_spec
    spec = random_spec(seed)
    spec.set_env(vary_temp_enabled=0, temp_amplitude=0.0, volume_effects_enabled=0)
    spec.phase_offset = None
    return spec


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built (no reference tree on this box)")
@pytest.mark.parametrize("name", sorted(CASES) + ["random_a%d" % k for k in range(0, 24, 2)])
def test_oracle_cpu_mode_bit_equals_reference_steps(name):
    spec = passive_random_spec(int(name[8:])) if name.startswith("random_a") else make_spec(name)
    ref = RefSim(spec)
    orc = OracleSim(ref.desc, cpu_lib_mode=1)
    dt = float(np.float32(0.9 * ref.recommended_dt()))
    for chunk in range(4):
        assert ref.step(250, dt) == 250
        assert orc.step(250, dt) == 250
        sr, so = ref.state(), orc.state()
        assert_bit_equal(so, sr, KIN + LINKF + LINKS + ["link_strain", "link_max_strain", "link_stress", "vox_flags"],
                         "%s after %d steps" % (name, 250 * (chunk + 1)))


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built (no reference tree on this box)")
def test_reference_demo_vxa_loads_and_matches_builder():
    """The reference's shipped demo (demos/basic/base.vxa) through its own reader vs the oracle stepping the exported model."""
    path = os.path.join(util.REFERENCE_TREE, "demos", "basic", "base.vxa")
    if not os.path.exists(path):
        pytest.skip("reference demo not present")
    lib = util.load_ref()
    h = lib.vxref_load_vxa(path.encode())
    assert lib.vxref_ok(h)
    d = lib.vxref_export(h)
    assert d.contents.n_voxels == 1 and d.contents.n_links == 0
    dt_rec = lib.vxref_recommended_dt(h)
    assert abs(dt_rec - 1.59155e-4) < 1e-8  # SURVEY.md finding 2
    orc = OracleSim(d, cpu_lib_mode=1)
    dt = float(np.float32(0.9 * dt_rec))
    lib.vxref_step(h, 2000, dt)
    orc.step(2000, dt)
    sb = util.StateBuffers(1, 0)
    import ctypes as C
    assert lib.vxref_state(h, C.byref(sb.view)) == 0
    assert_bit_equal(orc.state(), sb.result(), KIN, "demo")


# ---------------------------------------------------------------- committed fixtures (travel to the GPU box)
def golden_path(name):
    return os.path.join(util.GOLDEN, name + ".json")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_cpu_mode_matches_golden_fixture(name):
    """Same check as above against the committed reference outputs (hex floats: exact)."""
    p = golden_path(name)
    assert os.path.exists(p), "missing fixture %s (run tests/golden/make_golden.py where /root/reference exists)" % p
    g = json.load(open(p))
    spec = make_spec(name)
    lib = util.load_engine()
    b, d = spec.build(lib)
    try:
        orc = OracleSim(d, cpu_lib_mode=1)
        dt = float.fromhex(g["dt"])
        assert orc.step(g["steps"], dt) == g["steps"]
        so = orc.state()
        for k, hexes in g["state"].items():
            want = np.array([float.fromhex(x) for x in hexes]).reshape(np.asarray(so[k]).shape)
            np.testing.assert_array_equal(np.asarray(so[k], np.float64), want, err_msg="%s.%s" % (name, k))
    finally:
        lib.vx3_builder_destroy(b)
