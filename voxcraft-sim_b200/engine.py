"""Python binding of the engine's C ABI (include/vx3_abi.h): one ``Batch`` = one device = one stream.

Plumbing only.  The CUDA library is required: construction raises if it is missing or no sm_100 device is
usable (there is no CPU fallback).
"""
import ctypes as C

from . import abi
from .libs import load_engine
from .model import StateBuffers


class EngineError(RuntimeError):
    pass


class Batch:
    def __init__(self, desc_ptrs, fma=False, device=0):
        self.lib = load_engine(fma)
        n = len(desc_ptrs)
        arr = (abi.ModelDesc * n)()
        for i, d in enumerate(desc_ptrs):
            C.memmove(C.byref(arr[i]), d, C.sizeof(abi.ModelDesc))
        self._arr = arr
        self.n = n
        self.h = C.c_void_p()
        rc = self.lib.vx3_batch_create(device, arr, n, C.byref(self.h))
        if rc != 0:
            raise EngineError("vx3_batch_create failed (%d): %s" % (rc, self.lib.vx3_last_error().decode()))
        self.sizes = [(d.contents.n_voxels, d.contents.n_links) for d in desc_ptrs]

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (what, rc, self.lib.vx3_last_error().decode()))

    def step(self, k, dt=None):
        if dt is None:
            self._check(self.lib.vx3_batch_step(self.h, k), "vx3_batch_step")
        else:
            self._check(self.lib.vx3_batch_step_dt(self.h, k, dt), "vx3_batch_step_dt")

    def run(self, max_steps=0, steps_per_launch=0, history=None):
        o = abi.RunOpts(max_steps, steps_per_launch, 1 if history is not None else 0)
        chunks = history

        def cb(user, sim, data, n):
            chunks.append((sim, C.string_at(data, n)))
        fn = abi.HISTORY_CB(cb) if history is not None else abi.HISTORY_CB()
        self._check(self.lib.vx3_batch_run(self.h, C.byref(o), fn, None), "vx3_batch_run")

    def sync(self):
        self._check(self.lib.vx3_batch_sync(self.h), "vx3_batch_sync")

    def state(self, sim=0, link_cap=None):
        nv, nl = self.sizes[sim]
        sb = StateBuffers(nv, link_cap or max(nl * 2 + 64, 64))
        self._check(self.lib.vx3_batch_state(self.h, sim, C.byref(sb.view)), "vx3_batch_state")
        return sb.result()

    def results(self):
        arr = (abi.Result * self.n)()
        self._check(self.lib.vx3_batch_results(self.h, arr), "vx3_batch_results")
        return list(arr)

    def positions(self, sim=0):
        """(init_pos, pos, matid) of one simulation, or of the whole batch concatenated when sim is None / -1."""
        import numpy as np
        if sim is None or sim < 0:
            sim, nv = -1, sum(s[0] for s in self.sizes)
        else:
            nv = self.sizes[sim][0]
        ip, p, m = np.empty((nv, 3)), np.empty((nv, 3)), np.empty(nv, np.int32)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self.lib.vx3_batch_positions(self.h, sim, dp(ip), dp(p), m.ctypes.data_as(C.POINTER(C.c_int32))), "vx3_batch_positions")
        return ip, p, m

    def recommended_dt(self, sim=0):
        v = C.c_double()
        self._check(self.lib.vx3_batch_recommended_dt(self.h, sim, C.byref(v)), "vx3_batch_recommended_dt")
        return v.value

    def timing(self):
        ms, n = C.c_double(), C.c_int64()
        self.lib.vx3_batch_last_timing(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def set_profiling(self, on=True, use_persistent=True):
        self._check(self.lib.vx3_batch_set_profiling(self.h, int(on), int(use_persistent)), "vx3_batch_set_profiling")

    def set_fused(self, on=True):
        self._check(self.lib.vx3_batch_set_fused(self.h, int(on)), "vx3_batch_set_fused")

    def fused_info(self):
        """(active, blocks, interior links, face links) of the fused step's block plan."""
        out = (C.c_int32 * 4)()
        self._check(self.lib.vx3_batch_fused_info(self.h, out), "vx3_batch_fused_info")
        return tuple(out)

    def kernel_stats(self):
        out, i = {}, 0
        while True:
            name = C.create_string_buffer(32)
            ms, n = C.c_double(), C.c_int64()
            if self.lib.vx3_batch_kernel_stats(self.h, i, name, 32, C.byref(ms), C.byref(n)) != 0:
                break
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out

    # ---- slab decomposition of one body (include/vx3_abi.h vx3_batch_halo_*) ----
    def halo_setup(self, side, send_vox, recv_vox):
        import numpy as np
        sv = np.ascontiguousarray(send_vox, dtype=np.int32)
        rv = np.ascontiguousarray(recv_vox, dtype=np.int32)
        self._check(self.lib.vx3_batch_halo_setup(self.h, side, len(sv), sv.ctypes.data_as(C.POINTER(C.c_int32)), len(rv),
                                                  rv.ctypes.data_as(C.POINTER(C.c_int32))), "vx3_batch_halo_setup")

    def halo_export(self, side):
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.vx3_batch_halo_export(self.h, side, C.cast(buf, C.c_void_p)), "vx3_batch_halo_export")
        return bytes(buf)

    def halo_connect(self, side, handle, peer_n_recv):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.lib.vx3_batch_halo_connect(self.h, side, C.cast(buf, C.c_void_p), peer_n_recv), "vx3_batch_halo_connect")

    def halo_connect_local(self, side, peer):
        self._check(self.lib.vx3_batch_halo_connect_local(self.h, side, peer.h), "vx3_batch_halo_connect_local")

    def step_async(self, k, dt=-1.0):
        self._check(self.lib.vx3_batch_step_async(self.h, k, dt), "vx3_batch_step_async")

    def com_sums(self, sim=0):
        out = (C.c_double * 6)()
        self._check(self.lib.vx3_batch_com_sums(self.h, sim, out), "vx3_batch_com_sums")
        return list(out)

    def counters(self, sim=0):
        out = (C.c_int64 * 8)()
        self._check(self.lib.vx3_batch_counters(self.h, sim, out), "vx3_batch_counters")
        return dict(attach=out[0], detach=out[1], links=out[2], cand_peak=out[4], fail_peak=out[5])

    def check_neighbor_search(self, sim=0, n_pairs=4096, seed=1):
        """(mismatches, positives) of the two-sided depth-5 neighbour search against the reference's path walk on random pairs."""
        mm, pos = C.c_int(0), C.c_int(0)
        self._check(self.lib.vx3_batch_check_neighbor_search(self.h, sim, n_pairs, seed, C.byref(mm), C.byref(pos)), "vx3_batch_check_neighbor_search")
        return mm.value, pos.value

    def close(self):
        if self.h:
            self.lib.vx3_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
