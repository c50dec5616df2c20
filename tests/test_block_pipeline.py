"""Host logic of the multi-GPU encoder (zdw_b200/host/block_pipeline.h): the windows cut on the host equal the ones the
sequential loop walks through, and blocks that finish out of order leave in file order.  CPU only."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_host_window_cuts_and_ordered_results(tmp_path):
    exe = tmp_path / "block_pipeline_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "zdw_b200" / "host"),
                    str(ROOT / "tests" / "block_pipeline_test.cpp"), "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
    assert p.stdout.strip().endswith("0 mismatches")
