"""The C ABI is usable from plain C: examples/c_driver.c is compiled with gcc against include/amg1d.h
(which must stay a C header) and linked with libamg1d.so - on CPU; run - on a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "agglomerationmultigrid1d_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_driver")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_driver.c"), "-L", LIBDIR, "-lamg1d", f"-Wl,-rpath,{LIBDIR}", "-lm",
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_driver_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libamg1d.so")):
        pytest.fail("libamg1d.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    _build(tmp_path)


@pytest.mark.gpu
def test_c_driver_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "8192"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
