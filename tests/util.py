"""Test plumbing: loads the CHECKERS (oracle restatement, reference CPU library) and provides small
model factories + comparison helpers.  Only tests/, smoke() and bench.py's cpu_baseline import this."""
import ctypes as C
import json
import os
import tempfile

import numpy as np

import __graft_entry__ as graft

graft.load_package()
from voxcraft_sim_b200 import abi  # noqa: E402
from voxcraft_sim_b200.libs import load_engine  # noqa: E402
from voxcraft_sim_b200.model import ModelSpec, StateBuffers  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libvx3_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvxref.so")
REF_OMP_SO = os.path.join(ROOT, "oracle", "_ref", "libvxref_omp.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_TREE = "/root/reference"

_libs = {}
P = C.POINTER


def load_oracle():
    if "oracle" not in _libs:
        lib = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        lib.vx3o_create.argtypes = [P(abi.ModelDesc), C.c_int]
        lib.vx3o_create.restype = vp
        lib.vx3o_destroy.argtypes = [vp]
        lib.vx3o_recommended_dt.argtypes = [vp]
        lib.vx3o_recommended_dt.restype = C.c_double
        lib.vx3o_step.argtypes = [vp, C.c_long, C.c_float]
        lib.vx3o_step.restype = C.c_long
        lib.vx3o_run.argtypes = [vp, C.c_long]
        lib.vx3o_run.restype = C.c_long
        lib.vx3o_result.argtypes = [vp, P(abi.Result), C.c_int]
        lib.vx3o_state.argtypes = [vp, P(abi.StateView)]
        lib.vx3o_counts.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_long), P(C.c_long)]
        lib.vx3o_surface.argtypes = [vp, P(C.c_int), C.c_int]
        lib.vx3o_eval.argtypes = [P(abi.Token), C.c_int, P(C.c_double)]
        lib.vx3o_eval.restype = C.c_double
        _libs["oracle"] = lib
    return _libs["oracle"]


def have_ref():
    return os.path.exists(REF_SO)


def load_ref(omp=False):
    key = "ref_omp" if omp else "ref"
    if key not in _libs:
        lib = C.CDLL(REF_OMP_SO if omp else REF_SO)
        vp = C.c_void_p
        lib.vxref_load_vxa.argtypes = [C.c_char_p]
        lib.vxref_load_vxa.restype = vp
        lib.vxref_ok.argtypes = [vp]
        lib.vxref_message.argtypes = [vp]
        lib.vxref_message.restype = C.c_char_p
        lib.vxref_export.argtypes = [vp]
        lib.vxref_export.restype = P(abi.ModelDesc)
        lib.vxref_recommended_dt.argtypes = [vp]
        lib.vxref_recommended_dt.restype = C.c_double
        lib.vxref_dtfrac.argtypes = [vp]
        lib.vxref_dtfrac.restype = C.c_double
        lib.vxref_step.argtypes = [vp, C.c_long, C.c_float]
        lib.vxref_step.restype = C.c_long
        lib.vxref_state.argtypes = [vp, P(abi.StateView)]
        lib.vxref_destroy.argtypes = [vp]
        _libs[key] = lib
    return _libs[key]


def ref_load_spec(spec, omp=False):
    """Write the spec as VXA and load it through the reference's own reader + Import."""
    lib = load_ref(omp)
    with tempfile.NamedTemporaryFile("w", suffix=".vxa", delete=False) as f:
        f.write(spec.to_vxa())
        path = f.name
    h = lib.vxref_load_vxa(path.encode())
    os.unlink(path)
    assert lib.vxref_ok(h), lib.vxref_message(h)
    return lib, h


# ------------------------------------------------------------------ model factories
from voxcraft_sim_b200.workloads import add_abc_materials, splitmix64  # noqa: E402


def cube_spec(n=(3, 3, 3), seed=42, actuated=True, lift=0, holes=0.0, name="cube", collisions=0, damping=(1.0, 0.8, 0.01)):
    """Multi-material (A/B/C) body of nx*ny*nz voxels, `lift` empty layers below, optional random holes."""
    nx, ny, nz = n
    spec = ModelSpec(0.01, name)
    add_abc_materials(spec)
    spec.set_env(bond_damping_z=damping[0], col_damping_z=damping[1], slow_damping_z=damping[2],
                 temp_enabled=1, vary_temp_enabled=int(actuated), temp_amplitude=20.0 if actuated else 0.0, temp_period=0.2)
    spec.set_options(enable_collision=collisions)
    r = splitmix64(seed)
    r2 = splitmix64(seed + 1)
    r3 = splitmix64(seed + 2)
    st = np.zeros((nz + lift, ny, nx), np.uint8)
    ph = np.zeros((nz + lift, ny, nx))
    for z in range(nz):
        for y in range(ny):
            for x in range(nx):
                m = 1 + r() % 3
                p = (r2() >> 11) / float(1 << 53)
                keep = ((r3() >> 11) / float(1 << 53)) >= holes
                if keep:
                    st[z + lift, y, x] = m
                ph[z + lift, y, x] = p
    spec.set_structure(st, phase_offset=ph if actuated else None)
    return spec


# ------------------------------------------------------------------ runners
class OracleSim:
    def __init__(self, desc_ptr, cpu_lib_mode=0):
        self.lib = load_oracle()
        self.h = self.lib.vx3o_create(desc_ptr, cpu_lib_mode)
        assert self.h
        self.nv = desc_ptr.contents.n_voxels
        self.nl = desc_ptr.contents.n_links

    def step(self, k, dt=-1.0):
        return self.lib.vx3o_step(self.h, k, dt)

    def run(self, max_steps=0):
        return self.lib.vx3o_run(self.h, max_steps)

    def recommended_dt(self):
        return self.lib.vx3o_recommended_dt(self.h)

    def counts(self):
        nv, nl, ns = C.c_int(), C.c_int(), C.c_int()
        a, d = C.c_long(), C.c_long()
        self.lib.vx3o_counts(self.h, nv, nl, ns, a, d)
        return dict(n_voxels=nv.value, n_links=nl.value, n_surface=ns.value, attach=a.value, detach=d.value)

    def state(self):
        c = self.counts()
        sb = StateBuffers(c["n_voxels"], c["n_links"])
        rc = self.lib.vx3o_state(self.h, C.byref(sb.view))
        assert rc == 0
        return sb.result()

    def result(self, refresh=True):
        r = abi.Result()
        self.lib.vx3o_result(self.h, C.byref(r), int(refresh))
        return r

    def close(self):
        if self.h:
            self.lib.vx3o_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefSim:
    def __init__(self, spec, omp=False):
        self.lib, self.h = ref_load_spec(spec, omp)
        self.desc = self.lib.vxref_export(self.h)
        self.nv, self.nl = self.desc.contents.n_voxels, self.desc.contents.n_links

    def step(self, k, dt):
        return self.lib.vxref_step(self.h, k, dt)

    def recommended_dt(self):
        return self.lib.vxref_recommended_dt(self.h)

    def state(self):
        sb = StateBuffers(self.nv, self.nl)
        assert self.lib.vxref_state(self.h, C.byref(sb.view)) == 0
        return sb.result()


from voxcraft_sim_b200.engine import Batch as EngineBatch  # noqa: E402  (the product binding; fails loudly without the CUDA library/GPU)


# ------------------------------------------------------------------ comparisons
KIN = ["pos", "orient", "lin_mom", "ang_mom"]
LINKF = ["link_force_neg", "link_force_pos", "link_moment_neg", "link_moment_pos"]
LINKS = ["link_pos2", "link_angle1v", "link_angle2v"]


def max_rel_err(a, b, scale=None):
    """max |a-b| / max(|b|_inf-per-array scale, tiny): a per-array relative error (robust for zeros)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    s = scale if scale is not None else max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def compare_states(sa, sb, keys, tol, what=""):
    worst = {}
    for k in keys:
        e = max_rel_err(sa[k], sb[k])
        worst[k] = e
        assert e <= tol, "%s: %s differs: rel err %.3e > %.1e" % (what, k, e, tol)
    return worst


def assert_bit_equal(sa, sb, keys, what=""):
    for k in keys:
        a, b = np.asarray(sa[k]), np.asarray(sb[k])
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        same = (a == b) | (np.isnan(a.astype(np.float64)) & np.isnan(b.astype(np.float64)))
        if not same.all():
            idx = np.argwhere(~same)[0]
            raise AssertionError("%s: %s not bit-equal at %s: %r vs %r (%d mismatches)" %
                                 (what, k, idx, a[tuple(idx)], b[tuple(idx)], (~same).sum()))


def desc_arrays(d):
    """Flatten a vx3_model_desc into numpy arrays / dicts for equality checks."""
    d = d.contents if hasattr(d, "contents") else d
    nv, nl = d.n_voxels, d.n_links

    def arr(p, n, dt):
        if not p or n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy()
    out = {"n_voxels": nv, "n_links": nl}
    out["ix"], out["iy"], out["iz"] = arr(d.ix, nv, np.int16), arr(d.iy, nv, np.int16), arr(d.iz, nv, np.int16)
    out["vox_mat"] = arr(d.vox_mat, nv, np.int32)
    out["pos"] = arr(d.pos, 3 * nv, np.float64)
    out["orient"] = arr(d.orient, 4 * nv, np.float64)
    out["vox_flags"] = arr(d.vox_flags, nv, np.int32)
    out["temp"] = arr(d.temp, nv, np.float32)
    out["phase_offset"] = arr(d.phase_offset, nv, np.float64)
    out["vox_links"] = arr(d.vox_links, 6 * nv, np.int32)
    for k in ("link_vneg", "link_vpos", "link_axis", "link_mat"):
        out[k] = arr(getattr(d, k), nl, np.int32)
    out["link_rest_length"] = arr(d.link_rest_length, nl, np.float64)
    out["link_transverse_area"] = arr(d.link_transverse_area, nl, np.float32)
    out["link_strain_ratio"] = arr(d.link_strain_ratio, nl, np.float32)

    def mat(m):
        r = {}
        for name, ct in abi.VoxelMaterial._fields_:
            if name in ("strain_data", "stress_data"):
                r[name] = [float(np.float32(getattr(m, name)[i])) for i in range(m.n_data)]
            elif name == "extScale":
                r[name] = list(m.extScale)
            else:
                r[name] = getattr(m, name)
        return r
    out["voxel_mats"] = [mat(d.voxel_mats[i]) for i in range(d.n_voxel_mats)]
    lm = []
    for i in range(d.n_link_mats):
        L = d.link_mats[i]
        r = mat(L.m)
        for name, _ in abi.LinkMaterial._fields_[1:]:
            r[name] = getattr(L, name)
        lm.append(r)
    out["link_mats"] = lm
    return out


def smoke_check():
    """__graft_entry__.smoke(): a small actuated body on cuda:0 through the C ABI vs the oracle."""
    spec = cube_spec((4, 4, 4), seed=7, actuated=True, name="smoke")
    lib = load_engine()
    b, d = spec.build(lib)
    try:
        eng = EngineBatch([d])
        orc = OracleSim(d)
        dt = np.float32(0.9 * orc.recommended_dt())
        eng.step(300, float(dt))
        orc.step(300, float(dt))
        se, so = eng.state(0), orc.state()
        worst = compare_states(se, so, KIN, 1e-9, "smoke")
        print("smoke ok: 4x4x4 actuated body, 300 steps, max rel err vs oracle:",
              {k: "%.2e" % v for k, v in worst.items()})
        eng.close()
    finally:
        lib.vx3_builder_destroy(b)
