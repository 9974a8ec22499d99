"""In-tree native build: nvcc for the sm_100a engine + g++ for the host model code.

Outputs (git-ignored, shipped to the GPU box by gpurun):
  voxcraft-sim_b200/lib/libvx3_b200.so      the product: compiled with -fmad=false so every fp64/fp32 operation rounds
                                            like the reference's x86-64 build (which does not contract) — the
                                            parity-grade build, and the one bench.py measures
  voxcraft-sim_b200/lib/libvx3_b200_fma.so  (only with `build.py --fma`) same sources with FMA contraction on (nvcc default): an
                                            experiment to measure what contraction would buy; NOT parity-grade,
                                            not built, loaded or tested by default
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")

ENGINE_CU = ["engine/vx3_engine.cu"]
HOST_CPP = ["host/vx3_materials.cpp", "host/vx3_builder.cpp", "host/vx3_xml.cpp", "host/vx3_vxa.cpp", "host/vx3_worker.cpp"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _all_deps():
    deps = []
    for base, _, files in os.walk(CSRC):
        deps += [os.path.join(base, f) for f in files]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return deps


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_lib(fma=False, force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    name = "libvx3_b200_fma.so" if fma else "libvx3_b200.so"
    out = os.path.join(LIBDIR, name)
    srcs = [os.path.join(CSRC, s) for s in ENGINE_CU + HOST_CPP if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _newer(out, _all_deps()):
        return out
    cmd = [_nvcc(), "-std=c++17", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-I", INCLUDE,
           "-I", CSRC] + ARCH
    if not fma:
        cmd += ["-fmad=false"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + srcs
    _run(cmd, verbose)
    return out


def build_exes(force=False, verbose=False):
    """vx3_node_worker and voxcraft-sim (drop-in executables) next to the library; rpath = $ORIGIN."""
    os.makedirs(LIBDIR, exist_ok=True)
    outs = []
    for name, src, link in (("vx3_node_worker", "exe/vx3_node_worker.cpp", True), ("voxcraft-sim", "exe/voxcraft-sim.cpp", False)):
        out = os.path.join(LIBDIR, name)
        srcp = os.path.join(CSRC, src)
        if force or _newer(out, [srcp, os.path.join(LIBDIR, "libvx3_b200.so")] + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]):
            cmd = ["g++", "-std=c++17", "-O2", "-I", INCLUDE, srcp, "-o", out]
            if link:
                cmd += ["-L", LIBDIR, "-lvx3_b200", "-Wl,-rpath,$ORIGIN"]
            _run(cmd, verbose)
        outs.append(out)
    return outs


def build_oracle(force=False, verbose=False):
    """Builds the CHECKERS (tests/bench cpu_baseline only): oracle restatement and, when the reference
    tree is present (this container), the unmodified reference CPU library under oracle/_ref."""
    odir = os.path.join(ROOT, "oracle")
    _run(["make", "-C", odir, "-j8", "oracle"], verbose)
    ref = os.environ.get("VX3_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "src", "old")):
        _run(["make", "-C", odir, "-j8", "ref", "REF=" + ref], verbose)
        try:  # multi-core CPU baseline (the reference's own USE_OMP path); optional
            _run(["make", "-C", odir, "-j8", "ref_omp", "REF=" + ref], verbose)
        except RuntimeError:
            print("note: OpenMP reference build unavailable (no libgomp for this g++)")
        # the reference's VX3 device sources compiled for the host (oracle/ref_vx3): pins the oracle's VX3-only behaviour
        _run(["make", "-C", odir, "-j8", "ref_vx3", "REF=" + ref], verbose)
    return os.path.join(odir, "libvx3_oracle.so")


def build_all(force=False, verbose=False, fma=False):
    a = build_lib(fma=False, force=force, verbose=verbose)
    build_exes(force=force, verbose=verbose)
    if fma:
        return a, build_lib(fma=True, force=force, verbose=verbose)
    return (a,)


if __name__ == "__main__":
    v = "-v" in sys.argv
    f = "-f" in sys.argv
    print(build_all(force=f, verbose=v, fma="--fma" in sys.argv))
    print(build_oracle(verbose=v))
