#!/usr/bin/env python
"""bench.py -- converged OS eigenvalues / second, Blasius temporal sweep (BASELINE.json).

One "step" = one pass of the hot path (batched CONTE_IC, shoot.f:391-813) over one
batch of synthetic sweep points: BASELINE config 5, the FIXED 4096 x 4096 (alpha, Re)
contebl grid, alpha_j = 0.05 + 0.25 j/4095, Re_i = 300 * 6^(i/4095) (contebl input
units), contebl defaults otherwise (nbl = nstep = 1000, CK45, testalpha = 10 deg,
y in [0, 17]); guesses are the bilinear interpolation of the committed 65 x 65
table bench_data/cfg5_seeds_65x65.npy.

Scaling: STRONG by default -- the one 4096 x 4096 grid is sharded over the N ranks
(cyclic 4096-point tiles of the flattened index, 16.7 M / N points per GPU), no
data-path collective, the per-point results gathered to rank 0 once per step inside
the timed region.  `--scaling weak` (and the `weak` key of the default run at N > 1)
gives every rank a full 4096 x 4096 batch instead (Re axis refined N-fold).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]
                  [--scaling strong|weak] [--no-cpu] [--no-e2e] [--no-secondary]

Prints ONE JSON line on rank 0 (see the task contract):
  value                device-resident throughput of the whole job
  e2e                  through the host-buffer C-ABI call (one process per GPU), copies inside the timed region
  e2e_single_process   (N > 1) rank 0 alone drives all N GPUs through ONE osb_create_multi context with pinned
                       host buffers -- the deployment of the single-threaded Fortran host -- and its results are
                       compared bit for bit with the gathered per-rank results
  roofline             FP64-pipe roofline of the shooting kernel (algorithmic flops / CUDA-event kernel time over
                       a DFMA micro-benchmark run in this process); traffic from the committed ncu capture
  cpu_baseline         the CPU oracle (a port of the reference; no Fortran compiler exists here) on the host
                       cores over a bounded sample of the same workload
  secondary            (N = 1) the other BASELINE configs, bounded to a few seconds each: 2 (conte channel grid),
                       3 (orrspace continuation lines), 4 (orrfdchan banded / orrcolchan dense, N = 64..512),
                       5b (orrncbl neutral points), K1 batched profiles -- each with value + roofline
  fp64_peak            DFMA and DMMA micro-benchmarks (also written to profiles/fp64_peak.json)
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
NRE = 4096
NSTEP = 1000
FLOPS_PER_STEP = 1122.0  # contebl step, algorithmic count
METRIC = "converged OS eigenvalues/sec (Blasius temporal sweep)"
UNIT = "eigenvalues/s"
TILE = 4096


# ----------------------------------------------------------------- workload
def _sweep_module():
    import importlib.util

    spec = importlib.util.spec_from_file_location("osb_sweep", ROOT / "os-stab_b200" / "sweep.py")
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)  # pure host logic; does not need the CUDA library
    return sweep


def workload_indices(world, rank, points=None, scaling="strong"):
    """Flattened (i_Re, j_alpha) indices owned by `rank` (cyclic tiles) and the Re-axis length."""
    nre = NRE if scaling == "strong" else NRE * world
    idx = _sweep_module().shard_indices(NA * nre, world, rank, TILE)
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


def cpu_run(npts_sample, cores):
    """The CPU oracle (port of the reference's CONTE_IC) over an evenly strided
    sample of the fixed 4096 x 4096 grid, one process per core.  Returns (eig/s, conv, passes, secs, n)."""
    from concurrent.futures import ProcessPoolExecutor

    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob

    ob.build()
    idx, nre = workload_indices(1, 0)
    # evenly spread over the flattened grid; the positions are deliberately not a
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
        v, conv, passes, dt, n = cpu_run(sample, cores)
        if s >= args.warmup:
            vals.append(v); ms.append(dt * 1e3)
    value = float(np.mean(vals)) if vals else 0.0
    sample_txt = (f"{n} points spread evenly over the fixed 4096x4096 grid of config 5, same seeds; {conv} converged, "
                  f"{passes / max(n, 1):.2f} passes/point")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ms)) if ms else None,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config 5: fixed 4096x4096 (alpha,Re) contebl grid, bounded CPU sample per step",
                   "nstep": NSTEP, "integrator": "RKCK45 fixed step", "seeds": "bilinear from 65x65 table"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------- ncu-derived traffic
def committed_traffic():
    """DRAM bytes per unit of work from the committed ncu captures (profiles/*_traffic.json, written by
    tools/ncu_traffic.py from `ncu --set full` raw pages): {kernel key: {"bytes_per_unit": .., "unit": .., "source": ..}}."""
    out = {}
    for p in sorted((ROOT / "profiles").glob("*_traffic.json")):
        try:
            out.update(json.loads(p.read_text()))
        except Exception:
            pass
    return out


# ----------------------------------------------------------------- secondary configs (N = 1)
def secondary_configs(eng, osb, peak_dfma, peak_dmma, hbm_gbs, traffic):
    """The other BASELINE configs, a few seconds each.  value + roofline per entry."""
    out = {}

    def best_of(fn, reps=2):
        fn()
        best, res = 1e30, None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = fn()
            best = min(best, time.perf_counter() - t0)
        return best, res

    def tr(key, units):
        t = traffic.get(key)
        return None if not t else float(t["bytes_per_unit"]) * units

    # ---- config 2: conte channel temporal grid, CK45, n = 2000
    try:
        prof = eng.channel_profile()
        opts = osb.shoot_defaults(osb.FLOW_CONTE)
        na = nr = 512
        alpha = 0.95 + 0.1 * np.arange(na) / (na - 1)
        Re = 9000.0 + 2000.0 * np.arange(nr) / (nr - 1)
        A, R = np.meshgrid(alpha, Re)
        ca = np.linspace(alpha[0], alpha[-1], 9); cr = np.linspace(Re[0], Re[-1], 9)
        CA, CR = np.meshgrid(ca, cr)
        coarse = eng.shoot_temporal(opts, prof, CA.ravel(), CR.ravel(), np.full(81, complex(0.23753, 0.00374)))
        cc = coarse["c"].reshape(9, 9)
        x = (A - alpha[0]) / (alpha[-1] - alpha[0]) * 8; y = (R - Re[0]) / (Re[-1] - Re[0]) * 8
        x0 = np.minimum(x.astype(int), 7); y0 = np.minimum(y.astype(int), 7); fx = x - x0; fy = y - y0
        guess = (1 - fy) * ((1 - fx) * cc[y0, x0] + fx * cc[y0, x0 + 1]) + fy * ((1 - fx) * cc[y0 + 1, x0] + fx * cc[y0 + 1, x0 + 1])
        dt, r = best_of(lambda: eng.shoot_temporal(opts, prof, A.ravel(), R.ravel(), guess.ravel()))
        prof.free()
        conv = int((r["status"] == 0).sum())
        tf = float(r["passes"].sum()) * 2000 * 990.0 / dt / 1e12  # conte step = 990 flop
        out["config2_conte_512x512"] = {
            "workload": "conte channel, 512x512 (alpha,Re) around (1, 1e4), n=2000, CK45, host buffers in/out",
            "value": conv / dt, "unit": UNIT, "converged": conv, "points": int(A.size), "passes_per_point": float(r["passes"].mean()),
            "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_dfma, "unit": "TFLOP/s", "frac": tf / peak_dfma,
                         "traffic": tr("shoot_queue_kernel_conte", A.size)}}
    except Exception as e:  # noqa: BLE001
        out["config2_conte_512x512"] = {"error": repr(e)}

    # ---- config 3: orrspace continuation lines (sweep.inp as 4096 lines x 50 omega)
    try:
        fact = math.sqrt(2.0) / 1.7207876
        ymin, ymax = 0.0, 20.0 * fact
        prof = eng.solve_bl(osb.FLOW_ORRSPACE, osb.INT_RK4, 1000, ymin, ymax, 0.5, 0.0)
        opts = osb.shoot_defaults(osb.FLOW_ORRSPACE)
        opts.nstep, opts.ymin, opts.ymax, opts.testalpha = 1000, ymin, ymax, 0.9
        nl = 4096
        Re3 = (600.0 + 0.5 * np.arange(nl)) * fact
        om = np.full(nl, complex(0.042, 0.0)) * fact
        dom = np.full(nl, 0.002) * fact
        a_seed = complex(0.13809592, 0.0051958399) * fact
        i1000 = int(round((1000.0 - 600.0) / 0.5))
        a0 = np.zeros(nl, dtype=complex)
        for rng in (range(i1000, nl), range(i1000 - 1, -1, -1)):  # first-point guesses: march in Re from the deck's point
            cur = a_seed if rng.start == i1000 else a0[i1000]
            for lo in range(rng.start, rng.stop, rng.step * 64):
                idx = list(range(lo, max(min(lo + rng.step * 64, nl), -1), rng.step))
                res = eng.shoot_spatial_lines(opts, prof, Re3[idx], om[idx], dom[idx], np.full(len(idx), cur), 1)
                a0[idx] = res["alpha"][:, 0]
                cur = res["alpha"][-1, 0]
        dt, r = best_of(lambda: eng.shoot_spatial_lines(opts, prof, Re3, om, dom, a0, 50), reps=1)
        prof.free()
        conv = int((r["status"] == 0).sum())
        tf = float(r["passes"].sum()) * 1000 * 802.0 / dt / 1e12
        out["config3_orrspace_4096_lines"] = {
            "workload": "orrspace sweep.inp as 4096 continuation lines (Re_delta* = 600 + 0.5 i) x 50 omega, RK4",
            "value": conv / dt, "unit": "converged points/s", "converged": conv, "points": nl * 50,
            "passes_per_point": float(r["passes"].mean()),
            "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_dfma, "unit": "TFLOP/s", "frac": tf / peak_dfma,
                         "traffic": tr("shoot_pair_kernel_spatial", nl * 50),
                         "note": "bound by the critical path of the slowest continuation line (SURVEY H6: points of a line are sequential); 4096 lines cannot fill 148 SMs"}}
    except Exception as e:  # noqa: BLE001
        out["config3_orrspace_4096_lines"] = {"error": repr(e)}

    # ---- config 4: matrix path, 4096 alpha in [0.76, 1.16] at Re = 1e4, shifts interpolated from a 9-point scan
    for disc, label in ((osb.DISC_FD, "fd_banded"), (osb.DISC_CHEB, "cheb_dense")):
        for N in (64, 128, 256, 512):
            key = f"config4_{label}_N{N}"
            try:
                ch = eng.chan(disc, N)
                coarse = np.linspace(0.76, 1.16, 9)
                rc = ch.solve(coarse, 1.0e4)
                good = rc["omega"].real > 0
                npts = 4096
                alpha = np.linspace(0.76, 1.16, npts)
                sig = np.interp(alpha, coarse[good], rc["omega"].real[good]) + 1j * np.interp(alpha, coarse[good], rc["omega"].imag[good])
                dt, r = best_of(lambda: ch.solve(alpha, 1.0e4, sigma0=sig), reps=2)
                ch.free()
                n = N - 1
                conv = int((r["status"] == 0).sum())
                its = float(r["iters"].mean())
                e = {"workload": f"{'orrfdchan' if disc == osb.DISC_FD else 'orrcolchan'} operator N={N}, 4096 alpha, shifts given, host buffers in/out",
                     "value": conv / dt, "unit": UNIT, "converged": conv, "points": npts, "inverse_iterations_per_point": its}
                if disc == osb.DISC_FD:
                    # algorithmic bytes: factors (9+4) n 16 B written once and streamed once per
                    # inverse iteration, plus the iteration vectors (4 n 16 B per iteration)
                    bytes_inst = 13.0 * n * 16 * (1.0 + its) + 4.0 * n * 16 * its
                    gbs = bytes_inst * npts / dt / 1e9
                    e["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                                     "traffic": tr(f"chan_fd_banded_N{N}", npts)}
                else:
                    flops = ((8.0 / 3.0) * n ** 3 + 16.0 * n * n * its) * npts  # one LU per point (a re-shift adds one)
                    tf = flops / dt / 1e12
                    e["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peak_dmma, "unit": "TFLOP/s", "frac": tf / peak_dmma,
                                     "traffic": tr(f"chan_eig_dense_N{N}", npts),
                                     "note": "fp64 tensor (DMMA) peak from the in-process micro-benchmark; LU-equivalent flops at one factorisation per point"}
                out[key] = e
            except Exception as ex:  # noqa: BLE001
                out[key] = {"error": repr(ex)}

    # ---- config 5b: orrncbl neutral points
    try:
        ch = eng.chan(osb.DISC_CHEB_NUM, 64)
        Re5 = 6000.0 * (1.0 + np.arange(4096) / 1024.0)
        dt, r = best_of(lambda: ch.neutral(Re5, 1.0, sfac=0.01), reps=1)
        ch.free()
        conv = int((r["status"] == 0).sum())
        steps = float(r["steps"].mean())
        tf = conv * steps * 2.0 * (8.0 / 3.0) * 63 ** 3 / dt / 1e12  # >= 2 factorisations per Newton step (right + re-shift)
        out["config5b_orrncbl_neutral"] = {
            "workload": "orrncbl Newton search, N=64 collocation, 4096 Re in [6000, 30000), sfac=.01, alpha0=1",
            "value": conv / dt, "unit": "neutral points/s", "converged": conv, "points": 4096, "newton_steps_per_point": steps,
            "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_dfma, "unit": "TFLOP/s", "frac": tf / peak_dfma,
                         "traffic": tr("chan_neutral_warp_kernel", 4096),
                         "note": "latency-bound warp kernel (63x63 matrix in shared memory); LU flops lower bound"}}
    except Exception as e:  # noqa: BLE001
        out["config5b_orrncbl_neutral"] = {"error": repr(e)}

    # ---- self-seeding sweep driver: config-5 plane from ONE anchor guess (no oracle-made table)
    try:
        ng = 1024
        al = (0.05 + 0.25 * np.arange(ng) / (ng - 1)) * S2
        Rg = 300.0 * 6.0 ** (np.arange(ng) / (ng - 1)) * S2
        prof = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
        opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
        t0 = time.perf_counter()
        r = eng.shoot_grid_temporal(opts, prof, al, Rg, 0.179 * S2, 580.0 * S2, complex(0.36413124, 0.79527387e-2), want_seed=True)
        dt = time.perf_counter() - t0
        prof.free()
        ok = r["status"] == 0
        d = np.abs(r["seed"] - r["c"])[ok]
        out["self_seeding_1024x1024"] = {
            "workload": "osb_shoot_grid_temporal: 1024 x 1024 points of the config-5 plane from the contebl default guess alone "
                        "(device-side continuation on a 65 x 65 coarse grid + bilinear seeds + one batched solve), host arrays out",
            "value": float(ok.sum()) / dt, "unit": UNIT, "seconds": dt, "converged": int(ok.sum()), "points": int(ok.size),
            "passes_per_point": float(r["passes"][ok].mean()), "coarse_phase": r["info"],
            "seed_error_fraction_below": {f"{t:g}": float(np.mean(d < t)) for t in (1e-6, 1e-5, 1e-4, 1e-3, 3e-3)},
            "seed_error_max": float(d.max())}
    except Exception as e:  # noqa: BLE001
        out["self_seeding_1024x1024"] = {"error": repr(e)}

    # ---- K1 batched: Falkner-Skan profile family
    try:
        nprof, nbl = 16384, 1000
        beta = np.linspace(-0.05, 0.3, nprof)
        f2p = 0.47 + 0.9 * beta
        dt, o = best_of(lambda: eng.solve_bl_batch(f2p, beta, osb.FLOW_CONTEBL, osb.INT_CK45, nbl, 0.0, 12.0), reps=1)
        gbs = nprof * 5.0 * (nbl + 1) * 8 / dt / 1e9  # algorithmic bytes: the five tables, once
        out["k1_blasius_batch"] = {
            "workload": "16384 Falkner-Skan profiles (beta in [-0.05, 0.3]) x 1000 steps through osb_profile_blasius_batch (host arrays out)",
            "value": nprof / dt, "unit": "profiles/s", "finite": bool(np.isfinite(o["U"]).all()),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                         "traffic": tr("blasius_kernel", nprof),
                         "note": "host call incl. D2H of the four tables; the march is sequential in y (latency-bound)"}}
    except Exception as e:  # noqa: BLE001
        out["k1_blasius_batch"] = {"error": repr(e)}
    return out


# ----------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def quiet_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on the
    first communicator when NCCL_DEBUG is set on the box): file descriptor 1 is pointed at stderr for the whole run and
    the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--points", type=int, default=None, help="per-rank points (default: the rank's share of the grid)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import os_stab_b200 as osb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # a host-side barrier for the single-process leg: ranks parked in an NCCL barrier keep a spinning kernel
        # on their GPU, which would time-slice against the kernels rank 0 launches there from its own context
        cpu_pg = dist.new_group(backend="gloo")

    eng = osb.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
    prof = eng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_pg)

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

    def timed_calls(reset, call, steps):
        """`steps` synchronous host calls, each preceded by an untimed reset of its in/out buffer; wall time of the
        calls alone, max over ranks"""
        barrier()
        tot = 0.0
        for _ in range(steps):
            reset()
            t0 = time.perf_counter()
            call()
            tot += time.perf_counter() - t0
        t = torch.tensor([tot * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0])

    class Leg:
        """One sharding of the workload resident on this rank's GPU + its step functions."""

        def __init__(self, scaling):
            self.scaling = scaling
            self.idx, self.nre = workload_indices(world, rank, args.points, scaling)
            alpha, Re, guess = build_inputs(self.idx, self.nre)
            n = self.n = self.idx.size
            # pinned host copies (the e2e arm copies from/to these every step)
            self.h_alpha = torch.from_numpy(alpha.view(np.float64).copy()).pin_memory()
            self.h_Re = torch.from_numpy(Re.copy()).pin_memory()
            self.h_guess = torch.from_numpy(guess.view(np.float64).copy()).pin_memory()
            self.h_c = torch.empty(2 * n, dtype=torch.float64).pin_memory()
            self.h_int = torch.empty(3 * n, dtype=torch.int32).pin_memory()
            with torch.cuda.stream(stream):
                self.d_alpha = self.h_alpha.to(dev, non_blocking=True)
                self.d_Re = self.h_Re.to(dev, non_blocking=True)
                self.d_guess = self.h_guess.to(dev, non_blocking=True)
                self.d_c = torch.empty_like(self.d_guess)
                self.d_int = torch.zeros(3 * n, dtype=torch.int32, device=dev)
                root = world > 1 and rank == 0
                self.gather_c = [torch.empty_like(self.d_c) for _ in range(world)] if root else None
                self.gather_i = [torch.empty_like(self.d_int) for _ in range(world)] if root else None
            stream.synchronize()

        def launch(self):
            n = self.n
            eng.shoot_temporal_dev(opts, prof, n, self.d_alpha.data_ptr(), self.d_Re.data_ptr(), self.d_c.data_ptr(),
                                   self.d_int.data_ptr(), self.d_int.data_ptr() + 4 * n, self.d_int.data_ptr() + 8 * n)

        def step_device(self):
            """inputs resident in HBM; result gathered to rank 0 (device)"""
            with torch.cuda.stream(stream):
                self.d_c.copy_(self.d_guess)
                self.launch()
                if world > 1:
                    dist.gather(self.d_c, self.gather_c, dst=0)
                    dist.gather(self.d_int, self.gather_i, dst=0)

        def reset_e2e(self):
            """c is in/out (guess -> result): restore the guesses; benchmark scaffolding, outside the timed call"""
            self.h_c.copy_(self.h_guess)

        def step_e2e(self):
            """the reference-facing call: HOST buffers in, HOST buffers out"""
            n = self.n
            c_np = self.h_c.numpy()
            i_np = self.h_int.numpy()
            eng._check(eng.lib.osb_shoot_temporal(
                eng.ctx, opts, prof.handle, n,
                osb.C.cast(self.h_alpha.data_ptr(), osb._dp), osb.C.cast(self.h_Re.data_ptr(), osb._dp),
                osb.C.cast(c_np.ctypes.data, osb._dp), osb.C.cast(i_np.ctypes.data, osb._ip),
                osb.C.cast(i_np.ctypes.data + 4 * n, osb._ip), osb.C.cast(i_np.ctypes.data + 8 * n, osb._ip),
                None, None, 0))

        def stats(self):
            """[converged, passes, points, maxcount, nonfinite] summed over ranks + this rank's (conv, passes)"""
            stream.synchronize()
            ints = self.d_int.cpu().numpy()
            n = self.n
            passes, status = ints[:n], ints[2 * n:]
            st = torch.tensor([float((status == 0).sum()), float(passes.sum()), float(n),
                               float((status == 1).sum()), float((status == 2).sum())], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(st, op=dist.ReduceOp.SUM)
            return [float(x) for x in st], float((status == 0).sum()), float(passes.sum())

        def free(self):
            for k in list(self.__dict__):
                if k.startswith(("h_", "d_", "gather_")):
                    setattr(self, k, None)

    leg = Leg(args.scaling)
    n = leg.n

    # ---- FP64 peaks (roofline denominators), measured live on this rank's GPU
    peak_tf, peak_mhz = eng.measure_fp64_peak(150.0)
    peak_dmma = eng.measure_dmma_peak(150.0)

    # ---- warm-up, then the timed device-resident region
    for _ in range(max(args.warmup, 0)):
        leg.step_device()
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms = timed(leg.step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - launches0

    # kernel-only time of one launch for the roofline (events tight around the call)
    barrier()
    k0 = torch.cuda.Event(enable_timing=True)
    k1 = torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(min(args.steps, 3)):
        with torch.cuda.stream(stream):
            leg.d_c.copy_(leg.d_guess)
            k0.record(stream)
            leg.launch()
            k1.record(stream)
        stream.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))
    (conv_total, passes_total, n_total, n_max, n_nan), my_conv, my_passes = leg.stats()

    # ---- e2e: host buffers through the C ABI, one process per GPU (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e_steps = min(args.steps, 3)
        leg.reset_e2e()
        leg.step_e2e()
        e_wall = timed_calls(leg.reset_e2e, leg.step_e2e, e_steps)  # the host call is synchronous: wall clock of the calls
        e2e = {"value": conv_total * e_steps / (e_wall * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(n * 28), "steps": e_steps,
               "ms_per_step": e_wall / e_steps,
               "api": "osb_shoot_temporal (include/osb.h): pinned host buffers in/out, H2D+kernel+D2H per call, one process per GPU"}

    # ---- e2e_single_process (N > 1): ONE host process, ONE multi-device context, the whole grid; bit-identity check
    single = None
    if world > 1 and not args.no_e2e and args.scaling == "strong" and args.points is None:
        # full-order reference results from the gathered per-rank shards (rank 0)
        full = None
        if rank == 0:
            sweep = _sweep_module()
            npts = NA * NRE
            full_c = np.empty(2 * npts); full_i = np.empty((3, npts), dtype=np.int32)
            for r in range(world):
                ridx = sweep.shard_indices(npts, world, r, TILE)
                gc = leg.gather_c[r].cpu().numpy().reshape(-1, 2)
                gi = leg.gather_i[r].cpu().numpy().reshape(3, -1)
                full_c.reshape(-1, 2)[ridx] = gc
                full_i[:, ridx] = gi
            full = (full_c, full_i)
        host_barrier()  # every other rank idles on the HOST from here: rank 0 drives their GPUs through its own context
        if rank == 0:
            try:
                meng = osb.Engine(devices=list(range(world)))
                mprof = meng.solve_bl(osb.FLOW_CONTEBL, osb.INT_CK45, 1000, 0.0, 17.0, 0.5, 0.0)
                idx_all = np.arange(NA * NRE, dtype=np.int64)
                a_all, r_all, g_all = build_inputs(idx_all, NRE)
                npts = idx_all.size
                hp_alpha = torch.from_numpy(a_all.view(np.float64).copy()).pin_memory()
                hp_Re = torch.from_numpy(r_all.copy()).pin_memory()
                hp_guess = torch.from_numpy(g_all.view(np.float64).copy()).pin_memory()
                hp_c = torch.empty(2 * npts, dtype=torch.float64).pin_memory()
                hp_int = torch.empty(3 * npts, dtype=torch.int32).pin_memory()

                def call():
                    meng._check(meng.lib.osb_shoot_temporal(
                        meng.ctx, opts, mprof.handle, npts,
                        osb.C.cast(hp_alpha.data_ptr(), osb._dp), osb.C.cast(hp_Re.data_ptr(), osb._dp),
                        osb.C.cast(hp_c.data_ptr(), osb._dp), osb.C.cast(hp_int.data_ptr(), osb._ip),
                        osb.C.cast(hp_int.data_ptr() + 4 * npts, osb._ip), osb.C.cast(hp_int.data_ptr() + 8 * npts, osb._ip),
                        None, None, 0))

                hp_c.copy_(hp_guess)
                call()
                s_steps = min(args.steps, 2)
                s_wall = 0.0
                for _ in range(s_steps):
                    hp_c.copy_(hp_guess)  # restore the guesses (c is in/out); not part of the timed call
                    t0 = time.perf_counter()
                    call()
                    s_wall += time.perf_counter() - t0
                s_wall /= s_steps
                ci = hp_int.numpy().reshape(3, -1)
                same = (np.array_equal(hp_c.numpy().view(np.int64), full[0].view(np.int64))
                        and np.array_equal(ci, full[1]))
                single = {"value": float((ci[2] == 0).sum()) / s_wall, "unit": UNIT, "ms_per_step": s_wall * 1e3, "steps": s_steps,
                          "h2d_bytes_per_step": int(npts * 40), "d2h_bytes_per_step": int(npts * 28), "devices": world,
                          "bit_identical_to_per_rank_results": bool(same),
                          "api": "osb_create_multi(ndev=N) + osb_shoot_temporal from ONE host thread (the Fortran host's deployment): "
                                 "cyclic tiles dealt to the devices inside the library, pinned host buffers in/out"}
                del hp_alpha, hp_Re, hp_guess, hp_c, hp_int
                mprof.free()
                meng.close()
            except Exception as e:  # noqa: BLE001
                single = {"error": repr(e)}
        host_barrier()

    # ---- the other scaling mode as an extra key (N > 1): weak next to the strong headline
    other = None
    if world > 1 and args.points is None:
        other_mode = "weak" if args.scaling == "strong" else "strong"
        leg.free()
        import gc
        gc.collect()
        leg2 = Leg(other_mode)
        leg2.step_device()
        o_steps = min(args.steps, 2)
        o_ms, _ = timed(leg2.step_device, o_steps)
        (o_conv, o_passes, o_n, _, _), _, _ = leg2.stats()
        other = {"scaling": other_mode, "value": o_conv * o_steps / (o_ms * 1e-3), "unit": UNIT, "ms_per_step": o_ms / o_steps,
                 "steps": o_steps, "points": int(o_n), "points_per_gpu": leg2.n}
        leg2.free()

    # ---- CPU baseline + secondary configs + peak record (rank 0, N = 1 only)
    cpu = None
    secondary = None
    hbm_gbs, hbm_src = 6552.0, "fallback: B200_PROFILING.md"
    try:
        mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        hbm_gbs, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    traffic = committed_traffic()
    if rank == 0 and world == 1:
        if not args.no_secondary and args.points is None:
            leg.free()
            secondary = secondary_configs(eng, osb, peak_tf, peak_dmma, hbm_gbs, traffic)
        if not args.no_cpu:
            cores = host_cores()
            v, conv, cpasses, dt, ns = cpu_run(256 * cores, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{ns} points spread evenly over the same 4096x4096 grid, same seeds, one process per core; "
                             f"{conv} converged, {cpasses / max(ns, 1):.2f} passes/point, {dt:.1f} s wall",
                   "per_core": v / cores}

    if rank == 0:
        peaks = {"dfma_tflops": peak_tf, "dmma_tflops": peak_dmma, "sm_mhz_est_from_dfma": peak_mhz,
                 "theoretical_dfma_tflops_at_max_clock": 148 * 64 * 2 * 1.965e9 / 1e12,
                 "clocks": clocks, "how": "osb_measure_fp64_peak / osb_measure_dmma_peak: 8 independent DFMA chains per thread / "
                                          "8 independent mma.sync.m8n8k4.f64 chains per warp, no memory traffic, best of 5"}
        try:
            for d in (ROOT / "profiles", ROOT / "gpurun_out"):
                if d.is_dir():
                    (d / "fp64_peak.json").write_text(json.dumps(peaks, indent=1) + "\n")
        except OSError:
            pass
        value = conv_total * args.steps / (dev_ms * 1e-3)
        flops_launch = my_passes * NSTEP * FLOPS_PER_STEP
        achieved = flops_launch / (kernel_ms * 1e-3) / 1e12
        tkey = traffic.get("shoot_queue_kernel")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"config 5: 4096 x {leg.nre} (alpha,Re) contebl grid, {int(n_total)} points "
                            f"({n} per GPU, cyclic 4096-point tiles), alpha 0.05..0.30, Re 300..1800",
                "nbl": 1000, "nstep": NSTEP, "integrator": "RKCK45 fixed step (USE_RKCK45)", "testalpha_deg": 10.0,
                "seeds": "bilinear interpolation of bench_data/cfg5_seeds_65x65.npy",
                "l2": f"inputs (40 B/point, {n * 40 / 1e6:.0f} MB per GPU) "
                      + ("exceed the 126 MB L2; no flush needed" if n * 40 > 126e6 else "are re-copied from the guess buffer every step"),
                "points_per_gpu": n,
            },
            "converged": int(conv_total), "points": int(n_total), "maxcount": int(n_max), "nonfinite": int(n_nan),
            "passes_per_point": passes_total / max(n_total, 1.0),
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks,
            "e2e": e2e,
            "e2e_single_process": single,
            "other_scaling": other,
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None,
                "executed_pipe_active_ncu": (tkey or {}).get("fp64_pipe_active_pct"),
                "traffic": float(tkey["bytes_per_unit"]) * n if tkey else None,
                "traffic_unit": "bytes per launch (dram read+write), per-point figure of the committed ncu capture x points of this launch",
                "traffic_source": (tkey or {}).get("source"),
                "algorithmic_bytes": float(n) * 68.0,
                "kernel": "osb::shoot_queue_kernel<CK45, Hermitian>" if not os.environ.get("OSB_NO_QUEUE") else "osb::shoot_kernel<CK45, Hermitian>",
                "kernel_ms": kernel_ms,
                "flops_per_launch": flops_launch,
                "how": "algorithmic flops = sum(passes) x nstep x 1122 (SURVEY 8d; ~920 are executed, the rest is spline/acos work the "
                       "kernel eliminated -- ncu's pipe-active figure is printed beside it) / CUDA-event time of the launch on the "
                       "library's stream; peak = DFMA micro-benchmark (osb_measure_fp64_peak) run in this process "
                       f"(est. {peak_mhz:.0f} MHz at 64 DFMA/clk/SM); MEASURED_PEAKS.json carries no fp64 figure",
            },
            "fp64_peak": {"dfma_tflops": peak_tf, "dmma_tflops": peak_dmma, "hbm_gbs": hbm_gbs, "hbm_source": hbm_src},
            "cpu_baseline": cpu,
            "secondary": secondary,
        }
        emit(line)
    barrier()
    # torch frees pinned/device buffers by recording events on the library's stream:
    # release them before the stream goes away with the context
    leg.free()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    prof.free()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
