#!/usr/bin/env python
"""Digest the `ncu --set full` reports of tools/capture_profiles.sh (gpurun_out/r02_<case>.ncu-rep) into
  profiles/r02_<case>_ncu_raw.csv   the raw-page metrics that matter (duration, DRAM bytes, pipes, occupancy, stalls)
  profiles/r02_traffic.json         DRAM bytes per unit of work per kernel family -- what bench.py reports as
                                    roofline.traffic (x the units of its own launch)
Run here (no GPU needed): python tools/ncu_traffic.py"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

# case -> (key in r02_traffic.json, unit name, units per captured launch)
CASES = {
    "shoot": ("shoot_queue_kernel", "point", 524288),
    "conte": ("shoot_queue_kernel_conte", "point", 262144),
    "spatial": ("shoot_pair_kernel_spatial", "point", 4096 * 50),
    "dense512": ("chan_eig_dense_N512", "instance", 296),
    "dense256": ("chan_eig_dense_N256", "instance", 1184),
    "dense128": ("chan_eig_dense_N128", "instance", 4096),
    "dense64": ("chan_eig_dense_N64", "instance", 4096),
    "band512": ("chan_fd_banded_N512_32768", "instance", 32768),
    "band512_small": ("chan_fd_banded_N512", "instance", 4096),
    "band64": ("chan_fd_banded_N64", "instance", 32768),
    "neutral": ("chan_neutral_warp_kernel", "neutral point", 4096),
    "blasius": ("blasius_kernel", "profile", 16384),
}
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum")
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3,
              "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def raw_page(rep):
    if rep.suffix == ".csv":
        txt = rep.read_text()
    else:
        txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def main():
    tj = ROOT / "profiles" / "r02_traffic.json"
    out = json.loads(tj.read_text()) if tj.exists() else {}  # cases not recaptured keep their entry
    for case, (key, unit, nunits) in CASES.items():
        rep = ROOT / "gpurun_out" / f"r02_{case}_rawpage.csv"
        if not rep.exists():
            rep = ROOT / "gpurun_out" / f"r02_{case}.ncu-rep"
        if not rep.exists():
            print("missing", case)
            continue
        m = raw_page(rep)
        if m is None:
            print("empty capture", case)
            continue
        kernel = m.get("Kernel Name", ("", "?"))[1]
        with open(ROOT / "profiles" / f"r02_{case}_ncu_raw.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            w.writerow(["Kernel Name", "", kernel])
            w.writerow(["units per launch (" + unit + ")", "", nunits])
            for h, (u, v) in m.items():
                if any(h == k or h.endswith(k) for k in KEEP) or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
                    w.writerow([h, u, v])

        def val(name):
            u, v = m[name]
            return float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)

        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        stalls = sorted(((float(v[1]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, v in m.items()
                         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")), reverse=True)
        tot = sum(s for s, _ in stalls) or 1.0
        out[key] = {
            "kernel": kernel, "unit": unit, "units_in_capture": nunits, "bytes_per_unit": (rd + wr) / nunits,
            "dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ms": val("gpu__time_duration.sum"),
            "fp64_pipe_active_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][1]),
            "dmma_pipe_active_pct": float(m["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"][1]),
            "warps_active_pct": float(m["sm__warps_active.avg.pct_of_peak_sustained_active"][1]),
            "registers_per_thread": int(float(m["launch__registers_per_thread"][1])),
            "top_stalls": [f"{n} {100 * s / tot:.0f}%" for s, n in stalls[:4]],
            "source": f"profiles/r02_{case}_ncu_raw.csv (ncu --set full --clock-control none, one launch, tools/capture_profiles.sh)",
        }
        print(f"{case:14s} {kernel[:60]:60s} {out[key]['duration_ms']:9.3f} ms  {out[key]['bytes_per_unit']:12.1f} B/{unit}  "
              f"fp64 {out[key]['fp64_pipe_active_pct']:5.1f}%  dmma {out[key]['dmma_pipe_active_pct']:5.1f}%  {out[key]['top_stalls']}")
    (ROOT / "profiles" / "r02_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    launches = ROOT / "gpurun_out" / "r02_launches.csv"
    if launches.exists():
        (ROOT / "profiles" / "r02_launches.csv").write_text(launches.read_text())


if __name__ == "__main__":
    main()
