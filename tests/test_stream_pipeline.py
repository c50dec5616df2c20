"""The read-ahead input window and the writer thread of convertDWfile (zdw_b200/host/stream_pipeline.h) against the
sequential window they replaced, compiled for the CPU with g++ and malloc as the allocator - no GPU involved."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_read_ahead_window_matches_sequential_window(tmp_path):
    exe = tmp_path / "stream_pipeline_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "zdw_b200" / "host"),
                    str(ROOT / "tests" / "stream_pipeline_test.cpp"), "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
    assert " 0 mismatches" in p.stdout


def test_decoder_sink_writes_in_order(tmp_path):
    """BufferedOutput::writeLater (unconvertDWfile's writer thread).  Links the host classes, hence libzdw_b200.so -
    it is only loaded, no CUDA call is made."""
    host = ROOT / "zdw_b200" / "host"
    lib = host / "build" / "libzdwhost.a"
    if not lib.exists() or not (ROOT / "zdw_b200" / "libzdw_b200.so").exists():
        import pytest
        pytest.skip("host classes not built (run __graft_entry__.build())")
    exe = tmp_path / "buffered_output_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(host), "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "buffered_output_test.cpp"), str(lib), "-L", str(ROOT / "zdw_b200"), "-lzdw_b200",
                    f"-Wl,-rpath,{ROOT / 'zdw_b200'}", "-o", str(exe)], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]
    assert p.stdout.strip().endswith("0 mismatches")
