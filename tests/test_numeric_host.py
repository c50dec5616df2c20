"""The numeric helpers shared by the kernels and the host (zdw_b200/csrc/common.cuh: strtoull semantics, CHAR tuples,
value widths, llutoa / lltoa text incl. the INT64_MIN quirk, the straight-line integer renderer) against libc, compiled
for the CPU with g++ - no GPU involved."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_numeric_helpers_match_libc(tmp_path):
    exe = tmp_path / "numeric_host"
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-I", str(ROOT / "zdw_b200" / "csrc"), "-I", "/usr/local/cuda/include",
                    str(ROOT / "tests" / "numeric_host.cpp"), "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "0 mismatches" in p.stdout
