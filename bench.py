#!/usr/bin/env python
"""bench.py -- converged OS eigenvalues / second, Blasius temporal sweep (BASELINE.json).

One "step" = one pass of the hot path (batched CONTE_IC, shoot.f:391-813) over one
batch of synthetic sweep points: BASELINE config 5, the 4096 x 4096 (alpha, Re)
contebl grid, alpha_j = 0.05 + 0.25 j/4095, Re_i = 300 * 6^(i/(NRe-1)) (contebl input
units), contebl defaults otherwise (nbl = nstep = 1000, CK45, testalpha = 10 deg,
y in [0, 17]); guesses are the bilinear interpolation of the committed 65 x 65
table bench_data/cfg5_seeds_65x65.npy.  Weak scaling: every rank owns a
4096 x 4096 batch; with N ranks the Re axis is refined to 4096*N values over the
same range and tiles of the flattened index are dealt cyclically to ranks.  No
data-path collective; the per-point results are gathered to rank 0 once per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]

Prints ONE JSON line on rank 0 (see the task contract): value = device-resident
throughput, e2e = through the host-buffer C-ABI call with copies inside the timed
region, roofline = FP64-pipe roofline of the shooting kernel, cpu_baseline = the
CPU oracle (a port of the reference; no Fortran compiler exists here) on the host
cores over a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

S2 = math.sqrt(2.0)
NA = 4096
NRE_PER_RANK = 4096
NSTEP = 1000
FLOPS_PER_STEP = 1122.0  # contebl step, algorithmic count
METRIC = "converged OS eigenvalues/sec (Blasius temporal sweep)"
UNIT = "eigenvalues/s"
TILE = 4096


# ----------------------------------------------------------------- workload
def workload_indices(world, rank, points=None):
    """Flattened (i_Re, j_alpha) indices owned by `rank` (cyclic tiles)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("osb_sweep", ROOT / "os-stab_b200" / "sweep.py")
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)  # pure host logic; does not need the CUDA library
    nre = NRE_PER_RANK * world
    idx = sweep.shard_indices(NA * nre, world, rank, TILE)
    if points is not None:
        idx = idx[: max(TILE, (points // TILE) * TILE)]
    return idx, nre


def build_inputs(idx, nre):
    """alpha, Re (solver units) and seeded guesses for the flattened indices."""
    seeds = np.load(ROOT / "bench_data" / "cfg5_seeds_65x65.npy")  # [i_Re, j_alpha]
    i = idx // NA
    j = idx % NA
    alpha_in = 0.05 + 0.25 * j / (NA - 1)
    Re_in = 300.0 * 6.0 ** (i / (nre - 1))
    # bilinear interpolation in (j, i) index space of the coarse table
    x = j * (64.0 / (NA - 1))
    y = i * (64.0 / (nre - 1))
    x0 = np.minimum(x.astype(np.int64), 63)
    y0 = np.minimum(y.astype(np.int64), 63)
    fx = x - x0
    fy = y - y0
    c = ((1 - fy) * ((1 - fx) * seeds[y0, x0] + fx * seeds[y0, x0 + 1])
         + fy * ((1 - fx) * seeds[y0 + 1, x0] + fx * seeds[y0 + 1, x0 + 1]))
    alpha = (alpha_in * S2).astype(np.complex128)  # contebl.f:181-182
    Re = Re_in * S2
    return alpha, Re, c.astype(np.complex128)


# ----------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "power_w_max": float(np.max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------- CPU arm
def host_cores():
    """Host threads this process may run on (affinity-aware; falls back to the CPU count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def _cpu_worker(args):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob

    alpha, Re, c = args
    o = ob.load(rebuild=False)
    prof, _, _ = o.solve_bl(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
    r = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, NSTEP, 10.0, 0.0, 17.0, prof, alpha, Re, c)
    return int((r["status"] == 0).sum()), int(r["passes"].sum())


def cpu_run(npts_sample, cores, world=1):
    """The CPU oracle (port of the reference's CONTE_IC) over an evenly strided
    sample of the rank-0 workload, one process per core.  Returns (eig/s, conv, passes, secs)."""
    from concurrent.futures import ProcessPoolExecutor

    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob

    ob.build()
    idx, nre = workload_indices(world, 0)
    # evenly spread over the flattened batch; the positions are deliberately not a
    # multiple of the row length so that both alpha and Re are sampled
    pos = np.unique(np.floor(np.arange(npts_sample) * ((idx.size - 1) / max(npts_sample - 1, 1))).astype(np.int64))
    sub = idx[pos]
    alpha, Re, c = build_inputs(sub, nre)
    parts = [(alpha[k::cores], Re[k::cores], c[k::cores]) for k in range(cores)]
    with ProcessPoolExecutor(cores) as ex:
        list(ex.map(_cpu_worker, [(p[0][:1], p[1][:1], p[2][:1]) for p in parts]))  # start workers, page in
        t0 = time.perf_counter()
        res = list(ex.map(_cpu_worker, parts))
        dt = time.perf_counter() - t0
    conv = sum(r[0] for r in res)
    passes = sum(r[1] for r in res)
    return conv / dt, conv, passes, dt, sub.size


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm for the path (oracle port;
    the Fortran cannot be compiled in this image) on all host cores."""
    if rank != 0:
        return
    cores = host_cores()
    sample = args.points or 256 * cores
    vals, ms = [], []
    conv = passes = n = 0
    for s in range(args.warmup + args.steps):
        v, conv, passes, dt, n = cpu_run(sample, cores, world)
        if s >= args.warmup:
            vals.append(v); ms.append(dt * 1e3)
    value = float(np.mean(vals)) if vals else 0.0
    sample_txt = f"{n} points spread evenly over rank 0's 4096x4096 batch of config 5, same seeds; {conv} converged, {passes / max(n, 1):.2f} passes/point"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)) if ms else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config 5: 4096x4096 (alpha,Re) contebl grid per GPU, bounded CPU sample per step",
                   "nstep": NSTEP, "integrator": "RKCK45 fixed step", "seeds": "bilinear from 65x65 table"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--points", type=int, default=None, help="per-rank points (default: the full 4096x4096 batch)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import os_stab_b200 as osb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    eng = osb.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    prof = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)

    idx, nre = workload_indices(world, rank, args.points)
    alpha, Re, guess = build_inputs(idx, nre)
    n = idx.size
    # pinned host copies (the e2e arm copies from/to these every step)
    h_alpha = torch.from_numpy(alpha.view(np.float64).copy()).pin_memory()
    h_Re = torch.from_numpy(Re.copy()).pin_memory()
    h_guess = torch.from_numpy(guess.view(np.float64).copy()).pin_memory()
    h_c = torch.empty(2 * n, dtype=torch.float64).pin_memory()
    h_int = torch.empty(3 * n, dtype=torch.int32).pin_memory()

    with torch.cuda.stream(stream):
        d_alpha = h_alpha.to(dev, non_blocking=True)
        d_Re = h_Re.to(dev, non_blocking=True)
        d_guess = h_guess.to(dev, non_blocking=True)
        d_c = torch.empty_like(d_guess)
        d_int = torch.zeros(3 * n, dtype=torch.int32, device=dev)
        gather_c = [torch.empty_like(d_c) for _ in range(world)] if (world > 1 and rank == 0) else None
        gather_i = [torch.empty_like(d_int) for _ in range(world)] if (world > 1 and rank == 0) else None
    stream.synchronize()

    def step_device():
        """inputs resident in HBM; result gathered to rank 0 (device)"""
        with torch.cuda.stream(stream):
            d_c.copy_(d_guess)
            eng.shoot_temporal_dev(opts, prof, n, d_alpha.data_ptr(), d_Re.data_ptr(), d_c.data_ptr(),
                                   d_int.data_ptr(), d_int.data_ptr() + 4 * n, d_int.data_ptr() + 8 * n)
            if world > 1:
                dist.gather(d_c, gather_c, dst=0)
                dist.gather(d_int, gather_i, dst=0)

    def step_e2e():
        """the reference-facing call: HOST buffers in, HOST buffers out"""
        h_c.copy_(h_guess)
        c_np = h_c.numpy()
        i_np = h_int.numpy()
        eng._check(eng.lib.osb_shoot_temporal(
            eng.ctx, opts, prof.handle, n,
            osb.C.cast(h_alpha.data_ptr(), osb._dp), osb.C.cast(h_Re.data_ptr(), osb._dp),
            osb.C.cast(c_np.ctypes.data, osb._dp), osb.C.cast(i_np.ctypes.data, osb._ip),
            osb.C.cast(i_np.ctypes.data + 4 * n, osb._ip), osb.C.cast(i_np.ctypes.data + 8 * n, osb._ip),
            None, None, 0))

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` calls between barriers; device time by CUDA events on the
        launching stream, max over ranks."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    # ---- FP64 peak (roofline denominator), measured live on rank 0's GPU
    peak_tf, peak_mhz = eng.measure_fp64_peak(150.0)

    # ---- warm-up, then the timed device-resident region
    for _ in range(max(args.warmup, 0)):
        step_device()
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - launches0

    # kernel-only time of one launch for the roofline (events tight around the call)
    barrier()
    k0 = torch.cuda.Event(enable_timing=True)
    k1 = torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(min(args.steps, 3)):
        with torch.cuda.stream(stream):
            d_c.copy_(d_guess)
            k0.record(stream)
            eng.shoot_temporal_dev(opts, prof, n, d_alpha.data_ptr(), d_Re.data_ptr(), d_c.data_ptr(),
                                   d_int.data_ptr(), d_int.data_ptr() + 4 * n, d_int.data_ptr() + 8 * n)
            k1.record(stream)
        stream.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))

    stream.synchronize()
    ints = d_int.cpu().numpy()
    passes, status = ints[:n], ints[2 * n:]
    stats = torch.tensor([float((status == 0).sum()), float(passes.sum()), float(n),
                          float((status == 1).sum()), float((status == 2).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    conv_total, passes_total, n_total, n_max, n_nan = [float(x) for x in stats]
    my_conv, my_passes = float((status == 0).sum()), float(passes.sum())

    # ---- e2e: host buffers through the C ABI (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e_steps = min(args.steps, 3)
        step_e2e()
        _, e_wall = timed(step_e2e, e_steps)  # the host call is synchronous: wall clock is the honest measure
        e2e = {"value": conv_total * e_steps / (e_wall * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 28), "steps": e_steps,
               "ms_per_step": e_wall / e_steps,
               "api": "osb_shoot_temporal (include/osb.h): pinned host buffers in/out, H2D+kernel+D2H per call"}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        v, conv, cpasses, dt, ns = cpu_run(256 * cores, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} points spread evenly over the same 4096x4096 batch, same seeds, one process per core; "
                         f"{conv} converged, {cpasses / max(ns, 1):.2f} passes/point, {dt:.1f} s wall",
               "per_core": v / cores}

    if rank == 0:
        value = conv_total * args.steps / (dev_ms * 1e-3)
        flops_launch = my_passes * NSTEP * FLOPS_PER_STEP
        achieved = flops_launch / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"config 5: 4096 x {nre} (alpha,Re) contebl grid, {int(n_total)} points "
                            f"({n} per GPU, cyclic 4096-point tiles), alpha 0.05..0.30, Re 300..1800",
                "nbl": 1000, "nstep": NSTEP, "integrator": "RKCK45 fixed step (USE_RKCK45)", "testalpha_deg": 10.0,
                "seeds": "bilinear interpolation of bench_data/cfg5_seeds_65x65.npy",
                "l2": "inputs (40 B/point, >600 MB per GPU) exceed the 126 MB L2; no flush needed",
                "points_per_gpu": n,
            },
            "converged": int(conv_total), "points": int(n_total), "maxcount": int(n_max), "nonfinite": int(n_nan),
            "passes_per_point": passes_total / max(n_total, 1.0),
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None,
                # ncu --set full, r01 capture (profiles/r01_shoot_queue_ncu_raw.csv): dram__bytes_read 40.2 B/point,
                # write-back of 28 B/point; scaled to this launch's point count
                "traffic": float(n) * (40.2 + 28.0), "traffic_unit": "bytes per launch (dram read+write)",
                "kernel": "osb::shoot_queue_kernel<CK45, Hermitian>" if not os.environ.get("OSB_NO_QUEUE") else "osb::shoot_kernel<CK45, Hermitian>",
                "kernel_ms": kernel_ms,
                "flops_per_launch": flops_launch,
                "how": "algorithmic flops = sum(passes) x nstep x 1122 (SURVEY 8d) / CUDA-event time of the launch on the "
                       "library's stream; peak = DFMA micro-benchmark (osb_measure_fp64_peak) run in this process "
                       f"(est. {peak_mhz:.0f} MHz at 64 DFMA/clk/SM); MEASURED_PEAKS.json carries no fp64 figure",
            },
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    barrier()
    # torch frees pinned/device buffers by recording events on the library's stream:
    # release them before the stream goes away with the context
    del h_alpha, h_Re, h_guess, h_c, h_int, d_alpha, d_Re, d_guess, d_c, d_int, gather_c, gather_i
    import gc
    gc.collect()
    torch.cuda.synchronize()
    prof.free()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
