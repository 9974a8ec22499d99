"""Locate and load the native libraries.  The engine library is REQUIRED: there is no CPU fallback —
if it is missing the import fails loudly instead of routing anywhere else."""
import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_cache = {}


def engine_path(fma=False):
    override = os.environ.get("VX3_ENGINE_LIB")  # developer experiments only (A/B builds of the same sources)
    if override:
        return override
    return os.path.join(HERE, "lib", "libvx3_b200_fma.so" if fma else "libvx3_b200.so")


def load_engine(fma=False):
    """fma=False: the product (parity-grade, -fmad=false).  fma=True: the FMA-contracted experimental build."""
    key = ("engine", fma)
    if key not in _cache:
        p = engine_path(fma)
        if not os.path.exists(p):
            raise ImportError("native engine %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % p)
        lib = C.CDLL(p)
        abi.declare_model_api(lib)
        abi.declare_engine_api(lib)
        _cache[key] = lib
    return _cache[key]
