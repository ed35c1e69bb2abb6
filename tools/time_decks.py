#!/usr/bin/env python
"""Wall time of the twin driver on the reference's FD decks (fd-64 ... fd-512: Re=1e4, alpha sweep,
no shift guesses: 48-shift scan + final solve per alpha), process start to exit."""
import subprocess
import tempfile
import time
from pathlib import Path

drv = Path(__file__).resolve().parent.parent / "os-stab_b200" / "host" / "osb_drivers"
for rep in range(2):
    for N, inc in ((64, "0.02"), (128, "0.01"), (256, "0.01"), (512, "0.01")):
        with tempfile.TemporaryDirectory() as d:
            deck = f"{N}\n10000\n0.76 1.16 {inc}\n0 0 1\neig-{N}\n0\n"
            t0 = time.perf_counter()
            r = subprocess.run([str(drv), "orrfdchan"], input=deck, capture_output=True, text=True, cwd=d)
            dt = time.perf_counter() - t0
            rows = sum(1 for ln in r.stdout.splitlines() if "E" in ln and ln.strip()[:1].isdigit())
            print(f"run {rep} N={N} inc={inc} rc={r.returncode} rows={rows} wall={dt:.3f} s", flush=True)
