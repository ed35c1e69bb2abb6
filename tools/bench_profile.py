#!/usr/bin/env python
"""Batched Blasius / Falkner-Skan profile generation (K1): nprof (beta, f2p) pairs.
Run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:blasius` for the kernel-only time and DRAM traffic (the host call also copies
and transposes the tables)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os_stab_b200 as osb  # noqa: E402

nprof = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nbl = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = osb.Engine(0)
beta = np.linspace(-0.05, 0.3, nprof)
f2p = 0.47 + 0.9 * beta
t0 = time.perf_counter()
out = eng.solve_bl_batch(f2p, beta, osb.FLOW_CONTEBL, osb.INT_CK45, nbl, 0.0, 12.0)
dt = time.perf_counter() - t0
print(f"nprof={nprof} nbl={nbl} host call {dt:.3f} s  passes {out['npass'].min()}..{out['npass'].max()}  "
      f"f''(0) range {out['f2p'].min():.6f}..{out['f2p'].max():.6f}  finite {np.isfinite(out['U']).all()}")
eng.close()
