#!/usr/bin/env python
"""Benchmark of the VX3 step loop on B200: voxel-steps/s (whole job, device-timed) + roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5] [--impl b200|reference]

One bench "step" = one pass of the hot path over one batch: `--sim-steps` calls of doTimeStep for every
simulation of the workload (default 1000; K*sim-steps = 100,000 for config 2 with the default K=100).
At N=1 the workload is BASELINE.json's config 2 (single 20x20x20 actuated body, collisions off).  For N>1 the
path shards by independent simulation: every rank steps its own copy of the per-GPU workload,
no data-path collective, scaling = weak; the only cross-rank traffic is the end-of-run gather of fitness results.

The same run also measures, in the same process group and with the same timing rules, the two multi-GPU workloads
north_star names, and reports them as sub-objects of the one JSON line (`--skip-extra` leaves them out; each timed region is
about a second long, so that the one nvidia-smi sample that may fall into it — every query stalls kernel launches for tens of
milliseconds — does not decide the number):
  "config3"  the vx3_node_worker batch: 512 random 10^3 robots PER GPU (4096 over 8), weak scaling, streaming kernels
  "config5"  ONE 200x200x100 body: undivided at N=1, cut into N x-slabs with halo exchange over peer memory at N>1
             (strong scaling), with a bit-exact self-check of the slab run against the undivided body on rank 0

Timing: W untimed warm-up steps, then exactly K steps between barrier+synchronize, device-timed with CUDA events
on the engine's own stream (vx3_batch_last_timing), max over ranks.  `value` starts with the model resident in
HBM; `e2e` re-creates the batch from HOST arrays every step (H2D), steps, and reads results + positions back
(D2H) through the C ABI, inside the timed region.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

METRIC = "voxel-steps/sec"
UNIT = "voxel-steps/s"


# stdout carries exactly ONE line, the JSON result: native libraries write banners to file descriptor 1 (NCCL prints its version
# there under NCCL_DEBUG=VERSION/WARN), so descriptor 1 is pointed at stderr for the whole run and the line goes to a saved copy
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def begin(self):
        """The timed region starts here: only samples from now on count.  The process is started before the warm-up and this
        call waits for its first sample: nvidia-smi takes a few hundred ms to come up, and while it initialises it holds
        driver locks that stall kernel launches — inside a short timed region that nearly doubled the measured step time."""
        if self.proc:
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:
                time.sleep(0.02)
        self.first = len(self.lines)

    def start(self):
        # one sampler per job (rank 0's GPU): every nvidia-smi query takes driver locks that stall kernel launches on the whole
        # box — eight samplers at 100 ms cost the launch-heavy config 3 run 13 % at 8 GPUs
        if env_rank()[0] != 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable" if env_rank()[0] == 0 else "sampled on rank 0"]}
        if len(self.lines) <= self.first:  # a timed region shorter than the sampling period: take the sample that ends it
            t0 = time.perf_counter()
            while len(self.lines) <= self.first and time.perf_counter() - t0 < 0.5:
                time.sleep(0.01)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, pkg, per_gpu_sims):
    from voxcraft_sim_b200 import workloads as W
    if name == "c2":
        specs = [W.c2_spec()]
        label = "config2: single 20x20x20 multi-material actuated body, collisions off"
    elif name == "c3":
        specs = [W.c3_spec(k) for k in range(per_gpu_sims)]
        label = "config3: batch of %d random 10x10x10 robots per GPU (vx3_node_worker fitness eval)" % per_gpu_sims
    elif name == "c4":
        specs = [W.c4_spec()]
        label = "config4: pile of 512 sticky 4x4x4 bodies, collisions + attach + detach"
    elif name == "c5":
        specs = [W.c5_spec()]
        label = "config5 (one GPU, no decomposition): single 200x200x100 body"
    else:
        raise SystemExit("unknown workload " + name)
    return specs, label


def reference_cpu_throughput(spec, target_seconds, omp=True):
    """The reference's own CPU implementation (src/old compiled unmodified into oracle/_ref) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util  # checker plumbing
    use_omp = omp and os.path.exists(util.REF_OMP_SO)
    if not os.path.exists(util.REF_SO) and not use_omp:
        return None
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if use_omp:  # torchrun presets OMP_NUM_THREADS=1: the CPU arm uses every host core it may run on
        os.environ["OMP_NUM_THREADS"] = str(cores)
    sim = util.RefSim(spec, omp=use_omp)
    dt = float(__import__("numpy").float32(0.9 * sim.recommended_dt()))
    sim.step(10, dt)
    t0 = time.perf_counter()
    sim.step(20, dt)
    per = (time.perf_counter() - t0) / 20
    chunk = max(20, int(1.0 / max(per, 1e-9)))
    t0 = time.perf_counter()
    done = 0
    while True:
        n = sim.step(chunk, dt)
        done += n
        el = time.perf_counter() - t0
        if el >= target_seconds or n < chunk:
            break
    return {"value": sim.nv * done / el, "unit": UNIT, "cores": cores if use_omp else 1, "kind": "reference",
            "sample": "%d doTimeStep calls of the same %d-voxel model through the reference CPU library (src/old, %s), %.1f s"
                      % (done, sim.nv, "-DUSE_OMP, %d threads" % cores if use_omp else "single thread as shipped", el),
            "seconds": el, "steps": done, "voxels": sim.nv}


def run_decomposed(args, lib, built, label, rank, local_rank, world, K, Wm, S, selfcheck_steps=20):
    """Config 5 on N GPUs: every rank builds the full model on the host, keeps its slab (+ ghost faces), wires the halo
    exchange with its neighbours and steps in lock step; value = voxels of the WHOLE body x steps / max device time.
    Returns the result line on rank 0 (None elsewhere)."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from voxcraft_sim_b200 import parallel
    from voxcraft_sim_b200.engine import Batch
    from voxcraft_sim_b200.workloads import alg_bytes_per_voxel_step
    _, d = built[0]
    nvox, nlinks = d.contents.n_voxels, d.contents.n_links
    dt = float(np.float32(d.contents.opt.dt_frac * lib.vx3_model_recommended_dt(d)))
    slab = parallel.partition_slabs(d, world, rank, axis=0)
    body = parallel.DecomposedBody(slab, dt, device=local_rank, fma=bool(args.fma))
    body.connect()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- self-check: the first steps of the slab run against the UNDIVIDED body stepped on rank 0, bit for bit ----
    selfcheck = None
    if selfcheck_steps > 0:
        body.step(selfcheck_steps)
        pos = body.batch.positions(0)[1]
        mine = hashlib.sha256(np.ascontiguousarray(pos[slab.owned]).tobytes()).hexdigest()
        table = [None] * world
        dist.all_gather_object(table, mine)
        if rank == 0:
            whole = Batch([d], device=local_rank)
            whole.set_profiling(False, use_persistent=False)
            whole.step(selfcheck_steps, dt)
            wpos = whole.positions(0)[1]
            whole.close()
            owner = parallel.slab_owner(d, world, axis=0)
            want = [hashlib.sha256(np.ascontiguousarray(wpos[owner == r]).tobytes()).hexdigest() for r in range(world)]
            selfcheck = {"steps": selfcheck_steps, "what": "sha256 of every slab's owned voxel positions == the same voxels of the undivided body stepped on rank 0",
                         "bit_exact": want == table}
        barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(Wm):
        body.step(S)
    barrier()
    sampler.begin()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(K):
        body.step(S)
        ms, nl = body.batch.timing()
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    com, sums = body.center_of_mass()  # the one collective of the path
    body.batch.set_profiling(True, use_persistent=False)
    prof_steps = min(S, 50)
    body.step(prof_steps)
    stats = {k: v for k, v in body.batch.kernel_stats().items() if v[1] > 0}
    body.batch.set_profiling(False, use_persistent=False)
    tt = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(launches), float(slab.owned.sum()), float(len(slab.voxels))], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        b_alg = alg_bytes_per_voxel_step(nvox, nlinks)
        dev_ms_max, wall_max = [float(x) for x in tt.tolist()]
        value = float(nvox) * S * K / (dev_ms_max * 1e-3)
        top = max(stats.items(), key=lambda kv: kv[1][0])
        n_l_local = body.batch.sizes[0][1]
        finfo = body.batch.fused_info()
        if top[0] == "k_fused":
            alg_launch = 228.0 * float(slab.owned.sum()) + 184.0 * finfo[2]
        elif top[0] == "k_links_face":
            alg_launch = 184.0 * finfo[3]
        else:
            alg_launch = 184.0 * n_l_local if top[0].startswith("k_links") else 228.0 * float(slab.owned.sum())
        avg_s = 1e-3 * top[1][0] / top[1][1]
        halo_ms = sum(v[0] for k, v in stats.items() if k.startswith("k_halo"))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dev_ms_max / K,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": label + ", cut into %d x-slabs with halo exchange over peer memory (NVLink)" % world, "voxels_total": nvox,
                           "links_total": nlinks, "voxels_owned_sum": int(cnt[1]), "voxels_held_sum_incl_ghosts": int(cnt[2]),
                           "face_voxels_rank0": {str(k): int(len(v)) for k, v in slab.send.items()}, "sim_steps_per_step": S,
                           "total_sim_steps": S * K, "build": "-fmad=false (parity-grade)", "path": ("fused blocks" if finfo[0] else "streaming") + " + k_halo",
                           "l2": "working set per GPU %.0f MB" % ((nvox * 228 + nlinks * 184) / world / 1e6),
                           "alg_bytes_per_voxel_step": b_alg, "wall_ms_per_step": 1e3 * wall_max / K, "center_of_mass": com},
                "selfcheck": selfcheck,
                # the exchange runs on a second stream under the interior link pass: what it costs is the part of the step that
                # the compute kernels do not account for (its own kernel time includes the wait for the neighbour)
                "compute_us_per_sim_step": 1e3 * sum(v[0] for k, v in stats.items() if not k.startswith("k_halo")) / prof_steps,
                "halo_exposed_us_per_sim_step": max(0.0, 1e3 * dev_ms_max / (K * S) - 1e3 * sum(v[0] for k, v in stats.items() if not k.startswith("k_halo")) / prof_steps),
                "k_halo_stream_us_per_sim_step": 1e3 * halo_ms / prof_steps,
                "roofline": {"bound": "hbm", "kernel": top[0], "achieved": alg_launch / avg_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg_launch / avg_s / 1e9 / peak, "traffic": None, "peak_source": peak_src, "rank": 0,
                             "kernel_ms": {k: round(v[0], 4) for k, v in stats.items()}, "kernel_launches": {k: v[1] for k, v in stats.items()}},
                "hbm_roofline_equiv_frac": value * b_alg / (peak * 1e9 * world), "cpu_baseline": None, "clocks": clocks,
                "gpu_launches": int(cnt[0]), "e2e": None}
    body.batch.close()
    return line


def measure_resident(batch, S, K, Wm, barrier, persistent=True):
    """W warm-up + K timed bench steps of S doTimeStep calls each on a resident batch; device time from the engine's own
    CUDA events; then one profiled pass for the per-kernel times.  Returns (dev_ms, launches, wall_s, kernel stats, prof_steps)."""
    for _ in range(Wm):
        batch.step(S)
    barrier()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(K):
        batch.step(S)
        ms, nl = batch.timing()
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    batch.set_profiling(True, use_persistent=persistent)
    prof_steps = min(S, 100)
    batch.step(prof_steps)
    stats = {k: v for k, v in batch.kernel_stats().items() if v[1] > 0}
    batch.set_profiling(False, use_persistent=persistent)
    return dev_ms, launches, wall, stats, prof_steps


def extra_resident(tag, specs, label, scaling, lib, rank, local_rank, world, barrier, S, K, Wm):
    """One of the extra workloads on a resident batch per rank (config 3 at any N, config 5 at N = 1): same timing rules as
    the headline (barrier + synchronize around exactly K steps, CUDA events, max over ranks, work summed over ranks)."""
    import torch
    import torch.distributed as dist
    from voxcraft_sim_b200.engine import Batch
    from voxcraft_sim_b200.workloads import alg_bytes_per_voxel_step
    built = [s.build(lib) for s in specs]
    descs = [d for _, d in built]
    nvox = sum(d.contents.n_voxels for d in descs)
    nlinks = sum(d.contents.n_links for d in descs)
    batch = Batch(descs, device=local_rank)
    dev_ms, launches, wall, stats, prof_steps = measure_resident(batch, S, K, Wm, barrier)
    res = batch.results()
    diverged = sum(1 for r in res if r.status == 2)
    counters = batch.counters(0) if tag == "c4" else None
    batch.close()
    for b, _ in built:
        lib.vx3_builder_destroy(b)
    tt = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(nvox) * S * K, float(launches), float(diverged)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    peak, _ = measured_peak_hbm()
    b_alg = alg_bytes_per_voxel_step(nvox, nlinks)
    dev_ms_max = float(tt[0])
    value = float(work[0]) / (dev_ms_max * 1e-3)
    total = sum(v[0] for v in stats.values()) or 1.0
    more = {}
    if counters is not None:  # which stretch of the pile's life was timed, and how much contact / attach / detach work it held
        more = {"timed_sim_steps": [Wm * S, (Wm + K) * S], "collision_counters_sim0": counters}
    return {**more, "workload": label, "value": value, "unit": UNIT, "n_gpus": world, "scaling": scaling, "steps": K, "warmup": Wm, "sim_steps_per_step": S,
            "ms_per_step": dev_ms_max / K, "us_per_sim_step": 1e3 * dev_ms_max / (K * S), "sims_per_gpu": len(descs), "voxels_per_gpu": nvox,
            "links_per_gpu": nlinks, "alg_bytes_per_voxel_step": b_alg, "hbm_roofline_equiv_frac": value * b_alg / (peak * 1e9 * world),
            "kernel_us_per_sim_step": {k: round(1e3 * v[0] / prof_steps, 3) for k, v in stats.items()},
            "kernel_share": {k: round(v[0] / total, 4) for k, v in stats.items()}, "gpu_launches": int(work[1]), "diverged_sims": int(work[2]),
            "wall_ms_per_step": 1e3 * float(tt[1]) / K}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sim-steps", type=int, default=1000, help="doTimeStep calls per bench step")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--sims-per-gpu", type=int, default=512, help="config 3: robots per GPU (4096 / 8)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fma", type=int, default=0, help="0 = the product (-fmad=false, parity-grade), 1 = FMA-contracted experimental build")
    ap.add_argument("--no-persistent", action="store_true", help="force the streaming kernels")
    ap.add_argument("--fused", action="store_true", help="opt into the fused block step (VX3_FUSED=1) instead of the two-pass streaming kernels")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="config 2 headline only: leave out the config 3 / 4 / 5 sub-measurements")
    args = ap.parse_args()
    claim_stdout()
    rank, local_rank, world = env_rank()
    K, Wm = args.steps, max(args.warmup, 0)
    graft.load_package()
    specs, label = build_workload(args.workload, None, args.sims_per_gpu)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        base = reference_cpu_throughput(specs[0], max(5.0, args.cpu_seconds))
        if base is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref is not built (reference sources are not on this box)"})
            return 0
        line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
                "ms_per_step": 1e3 * base["seconds"] / base["steps"] * args.sim_steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": label, "sim_steps_per_step": args.sim_steps, "note": "CPU reference (src/old) has no per-voxel phase actuation"},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import numpy as np
    import torch
    import torch.distributed as dist
    from voxcraft_sim_b200.engine import Batch
    from voxcraft_sim_b200.libs import load_engine
    from voxcraft_sim_b200.workloads import alg_bytes_per_voxel_step

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only (NCCL_DEBUG=VERSION prints a banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    fma = bool(args.fma)
    lib = load_engine(fma)
    built = [s.build(lib) for s in specs]
    descs = [d for _, d in built]
    nvox = sum(d.contents.n_voxels for d in descs)
    nlinks = sum(d.contents.n_links for d in descs)
    S = args.sim_steps

    if args.workload == "c5" and world > 1:
        # config 5 at N > 1: ONE body cut into slabs along x, one slab per GPU, halo exchange over peer memory inside the
        # step stream (strong scaling: the total work is fixed)
        line = run_decomposed(args, lib, built, label, rank, local_rank, world, K, Wm, S)
        if rank == 0:
            emit(line)
        for b, _ in built:
            lib.vx3_builder_destroy(b)
        dist.destroy_process_group()
        return 0

    if args.fused:
        os.environ["VX3_FUSED"] = "1"
    batch = Batch(descs, fma=fma, device=local_rank)
    if args.no_persistent:
        batch.set_profiling(False, use_persistent=False)
    finfo = batch.fused_info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(Wm):
        batch.step(S)
    barrier()
    sampler.begin()
    dev_ms, launches = 0.0, 0
    t_wall0 = time.perf_counter()
    for _ in range(K):
        batch.step(S)
        ms, nl = batch.timing()
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    res = batch.results()
    diverged = sum(1 for r in res if r.status == 2)
    counters = batch.counters(0) if args.workload == "c4" else None

    # ---- roofline of the dominant kernel: a profiled pass (CUDA events around every launch, same stream) ----
    batch.set_profiling(True, use_persistent=not args.no_persistent)
    prof_steps = min(S, 200)
    batch.step(prof_steps)
    stats = {k: v for k, v in batch.kernel_stats().items() if v[1] > 0}
    batch.set_profiling(False, use_persistent=not args.no_persistent)
    total_prof = sum(v[0] for v in stats.values()) or 1.0
    top = max(stats.items(), key=lambda kv: kv[1][0])
    peak, peak_src = measured_peak_hbm()
    b_alg = alg_bytes_per_voxel_step(nvox, nlinks)
    if top[0] == "k_persistent":
        # one launch covers many doTimeStep calls of the whole body: algorithmic bytes = B_alg * V * steps in the launch
        alg_bytes_launch = b_alg * nvox * prof_steps / top[1][1]
    elif top[0] == "k_links":
        alg_bytes_launch = 184.0 * nlinks
    elif top[0] == "k_voxels":
        alg_bytes_launch = 228.0 * nvox
    elif top[0] == "k_fused":  # all voxels + the links interior to a block (the face links run in k_links_face)
        alg_bytes_launch = 228.0 * nvox + 184.0 * batch.fused_info()[2]
    elif top[0] == "k_links_face":
        alg_bytes_launch = 184.0 * batch.fused_info()[3]
    else:
        alg_bytes_launch = b_alg * nvox
    avg_launch_s = 1e-3 * top[1][0] / top[1][1]
    achieved = alg_bytes_launch / avg_launch_s / 1e9
    traffic, traffic_src = None, None
    try:  # ncu-measured DRAM bytes per launch of this kernel on this workload (profiles/traffic.json), when a capture exists
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ent = json.load(f).get(args.workload, {}).get(top[0])
        if ent and (args.workload != "c3" or args.sims_per_gpu == 512):
            traffic, traffic_src = ent["bytes_per_launch"], ent["capture"] + (" (" + ent["note"] + ")" if ent.get("note") else "")
    except Exception:
        pass
    on_chip = top[0] == "k_persistent"
    roofline = {"bound": "latency (on-chip: state lives in registers / shared memory, DRAM traffic ~0; achieved / frac are the HBM-EQUIVALENT of the "
                         "algorithmic bytes, not HBM traffic)" if on_chip else "hbm",
                "equivalent": on_chip, "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_share_of_step": top[1][0] / total_prof,
                "alg_bytes_per_launch": alg_bytes_launch, "avg_launch_us": 1e6 * avg_launch_s,
                "kernel_ms": {k: round(v[0], 4) for k, v in stats.items()}, "kernel_launches": {k: v[1] for k, v in stats.items()}}

    # ---- e2e: host model -> vx3_batch_create (H2D) -> step -> results + positions (D2H), all inside the timed region ----
    e2e = None
    if not args.skip_e2e:
        e2e_steps = max(1, min(K, 10))
        h2d = 0
        for d in descs:
            m = d.contents
            h2d += m.n_voxels * (3 * 2 + 4 + 24 + 32 + 24 + 24 + 4 + 4 + 8 + 24 + 4) + m.n_links * (4 * 4 + 72 + 4 * 4 + 4 + 4 + 8 + 12)
        d2h = sum(C.sizeof(type(res[0])) for _ in res) + sum(d.contents.n_voxels * (48 + 4) for d in descs)
        def e2e_cycle():
            bt = Batch(descs, fma=fma, device=local_rank)
            if args.no_persistent:
                bt.set_profiling(False, use_persistent=False)
            bt.step(S)
            bt.results()
            bt.positions(None)  # every simulation's init_pos / pos / matid (SavePositionOfAllVoxels), one D2H
            bt.close()
        for _ in range(min(Wm, 3)):  # warm-up: the second arena + pinned staging buffer of the resource cache are allocated once
            e2e_cycle()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_cycle()
        barrier()
        e2e_t = time.perf_counter() - t0
        e2e = {"t": e2e_t, "steps": e2e_steps, "h2d": h2d, "d2h": d2h}

    # ---- aggregate over ranks: max time, summed work ----
    tt = torch.tensor([dev_ms, wall, e2e["t"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(nvox) * S * K, float(launches), float(diverged)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        # end-of-batch fitness gather (the only cross-device step of the path)
        fit = torch.tensor([r.fitness_score for r in res], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(fit) for _ in range(world)]
        dist.all_gather(gathered, fit)
    dev_ms_max, wall_max, e2e_max = [float(x) for x in tt.tolist()]
    total_work, total_launches, total_div = [float(x) for x in work.tolist()]

    # ---- the multi-GPU workloads north_star names, measured in the same run (sub-objects of the one line) ----
    extras = {}
    if args.workload == "c2" and not args.skip_extra:
        batch.close()
        batch = None
        from voxcraft_sim_b200 import workloads as W
        c3_specs = [W.c3_spec(k) for k in range(args.sims_per_gpu)]
        extras["config3"] = extra_resident("c3", c3_specs, "config3: batch of %d random 10x10x10 robots per GPU (vx3_node_worker fitness eval), %d in all"
                                           % (args.sims_per_gpu, args.sims_per_gpu * world), "weak", lib, rank, local_rank, world, barrier, 400, 20, 3)
        # config 4 from the start of the fall until the pile has settled and glued itself together (the contact phase costs most
        # once hundreds of bodies touch and every touching pair of one glued blob needs the depth-5 neighbour test)
        c4_specs, c4_label = build_workload("c4", None, 0)
        extras["config4"] = extra_resident("c4", c4_specs, c4_label + " (one pile per GPU)", "weak", lib, rank, local_rank, world, barrier, 1000, 7, 1)
        c5_specs, c5_label = build_workload("c5", None, 0)
        if world > 1:
            c5_built = [sp.build(lib) for sp in c5_specs]
            extras["config5"] = run_decomposed(args, lib, c5_built, c5_label, rank, local_rank, world, 20, 3, 100)
            for b5, _ in c5_built:
                lib.vx3_builder_destroy(b5)
        else:
            extras["config5"] = extra_resident("c5", c5_specs, c5_label, "strong", lib, rank, local_rank, world, barrier, 100, 6, 3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_baseline = reference_cpu_throughput(specs[0], args.cpu_seconds)
        if cpu_baseline:
            cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        value = total_work / (dev_ms_max * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": label, "voxels_per_gpu": nvox, "links_per_gpu": nlinks, "sims_per_gpu": len(descs),
                           "sim_steps_per_step": S, "total_sim_steps": S * K, "build": "fma" if fma else "-fmad=false (parity-grade)",
                           "path": "streaming" if args.no_persistent else "auto",
                           "fused_step": {"active": bool(finfo[0]), "blocks": finfo[1], "links_interior": finfo[2], "links_face_prepass": finfo[3]},
                           "l2": "state is mutated by every step (each step reads what the previous one wrote); working set "
                                 "%.1f MB %s the 126 MB L2" % ((nvox * 228 + nlinks * 184) / 1e6, "fits in" if nvox * 228 + nlinks * 184 < 100e6 else "exceeds"),
                           "alg_bytes_per_voxel_step": b_alg, "wall_ms_per_step": 1e3 * wall_max / K, "diverged_sims": int(total_div),
                           "collision_counters_sim0": counters, "timed_sim_steps": [Wm * S, (Wm + K) * S]},
                "roofline": roofline,
                "hbm_roofline_equiv_frac": value * b_alg / (peak * 1e9 * world),
                "cpu_baseline": cpu_baseline, "clocks": clocks, "gpu_launches": int(total_launches)}
        if e2e:
            ev = float(nvox) * S * e2e["steps"] * world / e2e_max
            line["e2e"] = {"value": ev, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": e2e["steps"], "what": "vx3_batch_create from host arrays + vx3_batch_step + vx3_batch_results/positions + destroy per step"}
        if total_div > 0:
            line["invalid"] = "%d simulation(s) diverged inside the timed region: their voxel-steps are counted but not computed" % int(total_div)
        for k, v in extras.items():
            if v is not None:
                line[k] = v
        emit(line)
    if batch is not None:
        batch.close()
    for b, _ in built:
        lib.vx3_builder_destroy(b)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
