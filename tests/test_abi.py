"""The C-ABI library loads and exports every symbol include/osb.h declares.
No compute calls here (no GPU in this container): with no device the library
must fail loudly instead of falling back to a CPU path."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "osb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(osb):
    lib = osb.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libosb.so does not export {n}"
    assert sorted(osb.EXPORTED_SYMBOLS) == names


def test_version_and_defaults(osb):
    lib = osb.load_library()
    assert lib.osb_version() == 2
    o = osb.shoot_defaults(osb.FLOW_CONTEBL, lib)
    # contebl.f:64-83 and shoot.f:461-463
    assert (o.nstep, o.maxcount, o.testalpha, o.ymin, o.ymax) == (1000, 20, 10.0, 0.0, 17.0)
    assert (o.errtol, o.eigtol, o.integrator) == (1e-14, 1e-14, osb.INT_CK45)
    o = osb.shoot_defaults(osb.FLOW_CONTE, lib)
    assert (o.nstep, o.errtol, o.ymin, o.ymax) == (2000, 1e-15, -1.0, 1.0)  # conte.f:61-68, shoot.f:53-55
    o = osb.shoot_defaults(osb.FLOW_ORRSPACE, lib)
    assert (o.maxcount, o.errtol, o.eigtol) == (50, 1e-8, 1e-12)  # orrspace.f:277-278
    bad = osb.ShootOpts()
    assert lib.osb_shoot_defaults(99, C.byref(bad)) == osb.E_ARG


def test_no_cpu_fallback_without_a_device(osb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(osb.OsbError) as ei:
        osb.Engine(0)
    assert ei.value.code == osb.E_CUDA


def test_product_package_does_not_touch_the_oracle():
    """Nothing under the product package may import, link or call the oracle."""
    banned = ("oracle/", "libos_oracle", "oracle_binding", "osr_")
    for p in (ROOT / "os-stab_b200").rglob("*"):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".f90", ".f"} or p.name == "Makefile":
            txt = p.read_text()
            for b in banned:
                assert b not in txt, f"{p} mentions {b}"


def test_shard_bounds_partition(osb):
    import numpy as np

    for npts, world, tile in ((10, 2, 4), (4096 * 7 + 5, 4, 4096), (3, 8, 4096), (0, 2, 16)):
        parts = [osb.shard_bounds(npts, world, r, tile) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(npts))


def test_header_is_plain_c_and_links(tmp_path):
    """include/osb.h is the drop-in boundary for a C / Fortran host: it must compile as C99 with no C++ or
    CUDA types in the signatures, and a C program must link against libosb.so.  Running it without a GPU
    must fail with OSB_E_CUDA, not crash (there is no CPU fallback)."""
    import subprocess

    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "osb.h"
int main(void) {
  osb_ctx *ctx = 0;
  osb_shoot_opts o;
  int rc = osb_shoot_defaults(OSB_FLOW_CONTEBL, &o);
  if (rc != 0 || o.nstep != 1000 || o.integrator != OSB_INT_CK45) return 10;
  double y[65];
  if (osb_chan_bl_points(OSB_DISC_CHEB, 64, 8.0, y) != 0 || !(y[64] == 0.0)) return 11;
  rc = osb_create(&ctx, 0);
  printf("osb_version %d create %d\n", osb_version(), rc);
  if (rc == 0) osb_destroy(ctx);
  return 0;
}
''')
    exe = tmp_path / "host"
    libdir = ROOT / "os-stab_b200"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                         "-L", str(libdir), "-losb", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "osb_version 2" in r.stdout
