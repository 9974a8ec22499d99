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
REF_VX3_SO = os.path.join(ROOT, "oracle", "_ref", "libvxref_vx3.so")  # the reference's VX3 device sources compiled for the host
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
        lib.vx3o_set_libm_jitter.argtypes = [C.c_ulonglong]
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


def have_ref_vx3():
    return os.path.exists(REF_VX3_SO)


def load_ref_vx3():
    if "ref_vx3" not in _libs:
        lib = C.CDLL(REF_VX3_SO)
        vp = C.c_void_p
        lib.vx3ref_create.argtypes = [C.c_char_p, P(abi.ModelDesc)]
        lib.vx3ref_create.restype = vp
        lib.vx3ref_ok.argtypes = [vp]
        lib.vx3ref_message.argtypes = [vp]
        lib.vx3ref_message.restype = C.c_char_p
        lib.vx3ref_init.argtypes = [vp]
        lib.vx3ref_recommended_dt.argtypes = [vp]
        lib.vx3ref_recommended_dt.restype = C.c_double
        lib.vx3ref_step.argtypes = [vp, C.c_long, C.c_float]
        lib.vx3ref_step.restype = C.c_long
        lib.vx3ref_run_simulation.argtypes = [vp]
        lib.vx3ref_run_simulation.restype = C.c_long
        lib.vx3ref_output.argtypes = [vp]
        lib.vx3ref_output.restype = C.c_char_p
        lib.vx3ref_eval.argtypes = [P(abi.Token), C.c_int, P(C.c_double)]
        lib.vx3ref_eval.restype = C.c_double
        lib.vx3ref_counts.argtypes = [vp, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int)]
        lib.vx3ref_surface.argtypes = [vp, P(C.c_int), C.c_int]
        lib.vx3ref_result.argtypes = [vp, P(abi.Result), C.c_int]
        lib.vx3ref_state.argtypes = [vp, P(abi.StateView)]
        lib.vx3ref_voxel_extras.argtypes = [vp, P(C.c_int32), P(C.c_int32)]
        _libs["ref_vx3"] = lib
    return _libs["ref_vx3"]


class Vx3RefSim:
    """The reference's own VX3 step loop (src/VX3/*.cu compiled for the host, oracle/ref_vx3) on a ModelSpec: the VXA text
    goes through CVX_Sim, the VX3-only settings come from the flat model `desc` (see vx3ref_harness.cpp)."""

    def __init__(self, spec, desc, vxa_text=None):
        self.lib = load_ref_vx3()
        with tempfile.NamedTemporaryFile("w", suffix=".vxa", delete=False) as f:
            f.write(vxa_text if vxa_text is not None else spec.to_vxa())
            path = f.name
        self.h = self.lib.vx3ref_create(path.encode(), desc)
        os.unlink(path)
        assert self.lib.vx3ref_ok(self.h), self.lib.vx3ref_message(self.h)

    def recommended_dt(self):
        return self.lib.vx3ref_recommended_dt(self.h)

    def step(self, k, dt=-1.0):
        return self.lib.vx3ref_step(self.h, k, dt)

    def run_simulation(self):
        """CUDA_Simulation itself; returns what it printed (history frames included)."""
        self.lib.vx3ref_run_simulation(self.h)
        return self.lib.vx3ref_output(self.h)

    def counts(self):
        nv, nl, ns, nm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.vx3ref_counts(self.h, nv, nl, ns, nm)
        return dict(n_voxels=nv.value, n_links=nl.value, n_surface=ns.value, n_link_mats=nm.value)

    def surface(self):
        c = self.counts()
        out = (C.c_int * max(c["n_surface"], 1))()
        n = self.lib.vx3ref_surface(self.h, out, c["n_surface"])
        return list(out[:n])

    def state(self):
        c = self.counts()
        sb = StateBuffers(c["n_voxels"], c["n_links"])
        assert self.lib.vx3ref_state(self.h, C.byref(sb.view)) == 0
        return sb.result()

    def result(self, refresh=True):
        r = abi.Result()
        self.lib.vx3ref_result(self.h, C.byref(r), int(refresh))
        return r

    def voxel_extras(self):
        n = self.counts()["n_voxels"]
        ea, rm = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.lib.vx3ref_voxel_extras(self.h, ea.ctypes.data_as(P(C.c_int32)), rm.ctypes.data_as(P(C.c_int32)))
        return ea, rm


# ------------------------------------------------------------------ model factories
from voxcraft_sim_b200.workloads import add_abc_materials, splitmix64  # noqa: E402
from scenarios import cube_spec  # noqa: E402,F401  (kept importable from util)


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


def libm_envelope(desc_ptr, steps, dt, checkpoints=None, seeds=(1, 2, 11, 23, 37), keys=None):
    """Element-wise spread of the oracle's state under a 1-ulp libm error model (vx3o_set_libm_jitter): the exact run plus
    len(seeds) replicas whose sin / cos / acos results are moved by one ulp: seed 1 = always up, seed 2 = always down (a
    consistently biased libm), other seeds = -1 / 0 / +1 at random per call.  Returns
    (exact_states, envelopes): one dict per checkpoint (steps, cumulative), envelope[k] = max over replicas of
    |replica - exact| (None where a replica's array shape differs, i.e. the topology itself is sensitive at the ulp level)."""
    lib = load_oracle()
    checkpoints = checkpoints or [steps]
    keys = keys or (KIN + LINKF + LINKS)

    def run(seed):
        lib.vx3o_set_libm_jitter(seed)
        try:
            o = OracleSim(desc_ptr)
            out, done = [], 0
            for c in checkpoints:
                assert o.step(c - done, dt) == c - done
                done = c
                out.append(o.state())
            return out
        finally:
            lib.vx3o_set_libm_jitter(0)
    exact = run(0)
    reps = [run(s) for s in seeds]
    envs = []
    for i, ex in enumerate(exact):
        env = {}
        for k in keys:
            a = np.asarray(ex[k], np.float64)
            e = np.zeros_like(a)
            for r in reps:
                b = np.asarray(r[i][k], np.float64)
                if b.shape != a.shape:
                    e = None
                    break
                e = np.maximum(e, np.abs(b - a))
            env[k] = e
        envs.append(env)
    return exact, envs


def link_force_floors(desc, so, rel=1e-9):
    """Absolute floors for the link force / moment arrays that are CONSISTENT with the position gate: a link force is stiffness x
    (difference of end positions), so two states whose positions agree to rel * max|pos| — the bar itself — can differ in a link
    force by k * rel * max|pos| (k = the stiffest link material's a1 or b1, N/m), and in a moment by that times the voxel size.
    Where this is larger than rel * max|force| (stiff links under small load) it is the honest floor."""
    d = desc.contents if hasattr(desc, "contents") else desc
    k = max([max(d.link_mats[i].a1, d.link_mats[i].b1) for i in range(d.n_link_mats)] + [0.0])
    for i in range(d.n_voxel_mats):  # attach-created links: (material, material) pairs may not be in the model's table
        k = max(k, float(d.voxel_mats[i].E) * float(d.voxel_mats[i].nomSize))
    size = max(float(d.voxel_mats[i].nomSize) for i in range(d.n_voxel_mats))
    dx = rel * float(np.max(np.abs(np.asarray(so["pos"], np.float64))))
    return {"link_force_neg": k * dx, "link_force_pos": k * dx, "link_moment_neg": k * dx * size, "link_moment_pos": k * dx * size}


def gate_within_envelope(se, so, env, keys, what="", factor=8.0, rel_floor=1e-9, abs_floor=None):
    """The tolerance gate of the GPU parity tests, element-wise:  |gpu - oracle| <= max(rel_floor * scale, factor * env)
    where scale = max |oracle| over the array (so zeros are handled) and env is the element's spread under the 1-ulp libm
    error model (libm_envelope; a handful of replicas, so an element's own spread is raised to the array's 90th percentile
    where that is larger).  rel_floor = 1e-9 is BASELINE.md's bar; the
    envelope term only ever loosens it where a 1-ulp difference in sin / cos / acos PROVABLY moves the oracle itself by
    more than that (momenta and link forces are differences of large terms).  abs_floor: per-key absolute floors
    (link_force_floors).  Returns the worst ratio err / tol per key."""
    worst = {}
    for k in keys:
        a, b = np.asarray(se[k], np.float64), np.asarray(so[k], np.float64)
        assert a.shape == b.shape, "%s: %s shape %s vs %s" % (what, k, a.shape, b.shape)
        if a.size == 0:
            continue
        scale = max(np.max(np.abs(b)), 1e-300)
        tol = np.full(a.shape, rel_floor * scale)
        if abs_floor is not None and k in abs_floor:
            tol = np.maximum(tol, abs_floor[k])
        if env is not None and env.get(k) is not None:
            e = env[k]
            tol = np.maximum(tol, factor * np.maximum(e, np.quantile(e, 0.9)))
        err = np.abs(a - b)
        ratio = float(np.max(err / tol))
        worst[k] = ratio
        if ratio > 1.0:
            i = np.unravel_index(np.argmax(err / tol), a.shape)
            raise AssertionError("%s: %s[%s] = %r vs oracle %r: |diff| %.3e > tol %.3e (1e-9 * scale = %.3e, libm envelope there %.3e)" %
                                 (what, k, i, a[i], b[i], err[i], tol[i], rel_floor * scale,
                                  0.0 if env is None or env.get(k) is None else env[k][i]))
    return worst


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
