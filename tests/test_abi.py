"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares, and the
ctypes mirror agrees with the compiled struct sizes.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import util
from voxcraft_sim_b200 import abi
from voxcraft_sim_b200.libs import engine_path, load_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vx3_[a-z0-9_]+)\s*\(", src)) - {"vx3_history_cb"})


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(engine_path())
    for header in ("vx3_abi.h", "vx3_model.h", "vx3_worker.h"):
        names = declared_functions(header)
        assert len(names) >= 3
        for n in names:
            assert hasattr(lib, n), "%s: symbol %s (declared in %s) is not exported" % (engine_path(), n, header)


def test_python_lists_match_headers():
    assert sorted(abi.ENGINE_SYMBOLS) == declared_functions("vx3_abi.h")
    assert sorted(abi.MODEL_SYMBOLS) == declared_functions("vx3_model.h")
    assert sorted(abi.WORKER_SYMBOLS) == declared_functions("vx3_worker.h")


def test_struct_sizes_match_the_compiled_headers():
    lib = load_engine()
    for name, cls in abi.STRUCTS.items():
        assert lib.vx3_abi_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.vx3_abi_sizeof(b"nope") == 0
    assert lib.vx3_abi_version() == 1


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device vx3_batch_create must return VX3_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        return
    lib = load_engine()
    spec = util.cube_spec((2, 2, 2))
    b, d = spec.build(lib)
    try:
        h = C.c_void_p()
        rc = lib.vx3_batch_create(0, d, 1, C.byref(h))
        assert rc == -3, rc
        assert b"no CPU fallback" in lib.vx3_last_error()
    finally:
        lib.vx3_builder_destroy(b)


def test_sort_results_fitness_descending_nan_last():
    lib = load_engine()
    vals = [0.5, float("nan"), 2.0, -1.0, float("nan"), 1.0]
    arr = (abi.Result * len(vals))()
    for i, v in enumerate(vals):
        arr[i].fitness_score = v
        arr[i].name = ("s%d" % i).encode()
    lib.vx3_sort_results(arr, len(vals))
    out = [arr[i].fitness_score for i in range(len(vals))]
    assert out[:4] == [2.0, 1.0, 0.5, -1.0]
    assert all(x != x for x in out[4:])
