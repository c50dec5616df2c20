#!/usr/bin/env python
"""cli_timing.py -- wall clock of the host tools (zdw_b200/bin/convertDWfile, unconvertDWfile) on a C4-shaped file of
several blocks, beside the compiled reference on one block of the same data.  What it is for: DESIGN.md section 5a -
the input read-ahead, the writer thread in front of the compressor and the two-context decode only show in the
wall clock of the tools, not in bench.py (which calls the C ABI directly).

    python tools/cli_timing.py [--blocks 8] [--rows 131072] [--compressors cat,gzip] [--dir /dev/shm]

Prints one JSON line per (tool, compressor).  `cat` = the pass-through stand-in of oracle/_ref/nocomp (the compressor
stage excluded, as in bench.py's reference arm); `gzip` = the real one.  Needs a GPU for our tools; the reference leg is
CPU only and is skipped with --no-reference."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (the synthetic C4 generator and the reference environment)

BIN = ROOT / "zdw_b200" / "bin"
REF = ROOT / "oracle" / "_ref"


def env_for(compressor: str):
    env = dict(os.environ)
    if compressor == "cat":
        env["PATH"] = f"{REF / 'nocomp'}:{env.get('PATH', '')}"
    return env


def timed(cmd, cwd, env):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed ({p.returncode}): {p.stderr[-400:].decode('latin1')}")
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--rows", type=int, default=bench.ROWS_PER_BLOCK)
    ap.add_argument("--compressors", default="cat,gzip")
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ap.add_argument("--block-bytes", type=int, default=0, help="--block-bytes for convertDWfile (0 = its default, 1 GiB)")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--gpus", default="", help="comma separated list of --gpus values to time, e.g. 1,2,4,8 (default: the tools' default)")
    ap.add_argument("--lanes", type=int, default=0, help="--lanes-per-gpu for our tools (0 = their default)")
    ap.add_argument("--stream-only", action="store_true", help="only the --stream leg")
    ap.add_argument("--stream", action="store_true",
                    help="also time `gzip -dc x.sql.gz | convertDWfile -i x.sql` (streaming input), with and without the read-ahead")
    args = ap.parse_args()

    synth = bench.Synth()
    work = Path(tempfile.mkdtemp(prefix="zdw_cli_", dir=args.dir))
    try:
        # ---- the input: `blocks` C4 blocks back to back in one file
        cap = synth.cap_for(args.rows)
        buf = (C.c_uint8 * cap)()
        src = work / "x.sql"
        one_block = 0
        with open(src, "wb") as f:
            for b in range(args.blocks):
                n = synth.block_into(b, args.rows, C.addressof(buf), cap)
                one_block = one_block or n
                f.write(memoryview(buf)[:n])
        (work / "x.desc.sql").write_bytes(synth.desc)
        tsv_bytes = src.stat().st_size
        extra = [f"--block-bytes={args.block_bytes}"] if args.block_bytes else []

        runs = [(comp, g) for comp in args.compressors.split(",") if comp for g in (args.gpus.split(",") if args.gpus else [""])]
        for comp, g in ([] if args.stream_only else runs):
            extra_g = ([f"--gpus={g}"] if g else []) + ([f"--lanes-per-gpu={args.lanes}"] if args.lanes else [])
            env = env_for(comp)
            for f in work.glob("x.zdw*"):
                f.unlink()
            t_enc = timed([str(BIN / "convertDWfile"), "-q", *extra, *extra_g, "x.sql"], work, env)
            zdw = work / "x.zdw.gz"
            out = work / "out"
            shutil.rmtree(out, ignore_errors=True)
            out.mkdir()
            t_dec = timed([str(BIN / "unconvertDWfile"), "-q", *extra_g, "-d", "out", "x.zdw.gz"], work, env)
            same = (out / "x.sql").stat().st_size == tsv_bytes and \
                subprocess.run(["cmp", "-s", str(out / "x.sql"), str(src)]).returncode == 0
            print(json.dumps({"tool": "zdw_b200", "compressor": comp, "gpus": g or "default", "blocks": args.blocks, "tsv_bytes": tsv_bytes,
                              "zdw_file_bytes": zdw.stat().st_size, "encode_s": round(t_enc, 3), "decode_s": round(t_dec, 3),
                              "encode_gbs": tsv_bytes / t_enc / 1e9, "decode_gbs": tsv_bytes / t_dec / 1e9,
                              "round_trip_identical": same}), flush=True)
            shutil.rmtree(out, ignore_errors=True)

        # ---- streaming input: the producer (gzip -dc) runs beside the encoder only if somebody keeps reading the pipe
        if args.stream or args.stream_only:
            sd = work / "stream"
            sd.mkdir()
            gz = shutil.which("gzip")  # the real one: env_for("cat") puts a pass-through `gzip` first on the tools' PATH
            subprocess.run(f"{gz} -1 -c {src} > {sd / 'x.sql.gz'}", shell=True, check=True)
            shutil.copy(work / "x.desc.sql", sd / "x.desc.sql")
            env = env_for("cat")
            os.symlink(src, sd / "x.sql")
            timed([str(BIN / "convertDWfile"), "-q", *extra, "x.sql"], sd, env)  # the file-input result to compare with
            os.rename(sd / "x.zdw.gz", sd / "file.zdw.gz")
            os.unlink(sd / "x.sql")
            t0 = time.perf_counter()
            subprocess.run(f"{gz} -dc {sd / 'x.sql.gz'} > /dev/null", shell=True, check=True)
            t_producer = time.perf_counter() - t0
            t0 = time.perf_counter()  # ... and through a pipe into a reader that does nothing else (what -i can reach at best)
            subprocess.run(f"{gz} -dc {sd / 'x.sql.gz'} | cat > /dev/null", shell=True, check=True)
            t_piped = time.perf_counter() - t0
            for label, extra_env in (("read-ahead", {}), ("sequential", {"ZDW_NO_READAHEAD": "1"})):
                for f in sd.glob("x.zdw*"):
                    f.unlink()
                e2 = dict(env)
                e2.update(extra_env)
                t0 = time.perf_counter()
                p = subprocess.run(f"{gz} -dc x.sql.gz | {BIN / 'convertDWfile'} -q -i {' '.join(extra)} x.sql", shell=True, cwd=sd, env=e2,
                                   capture_output=True)
                dt = time.perf_counter() - t0
                if p.returncode != 0:
                    raise RuntimeError(f"streaming convertDWfile failed ({p.returncode}): {p.stderr[-400:].decode('latin1')}")
                same = subprocess.run(["cmp", "-s", str(sd / "x.zdw.gz"), str(sd / "file.zdw.gz")]).returncode == 0
                print(json.dumps({"tool": "zdw_b200", "mode": "gzip -dc | convertDWfile -i", "input": label, "tsv_bytes": tsv_bytes,
                                  "encode_s": round(dt, 3), "producer_alone_s": round(t_producer, 3), "producer_into_cat_s": round(t_piped, 3),
                                  "encode_gbs": tsv_bytes / dt / 1e9, "same_zdw_as_file_input": same}), flush=True)

        # ---- the reference on the first block of the same file (single-threaded; the whole file would take minutes)
        if not args.no_reference and not args.stream_only and bench.have_ref():
            rd = work / "ref"
            rd.mkdir()
            with open(src, "rb") as f, open(rd / "x.sql", "wb") as g:
                g.write(f.read(one_block))
            t_enc, t_dec, nbytes = bench._ref_roundtrip(rd, rd / "x.sql", synth.desc)
            print(json.dumps({"tool": "reference", "compressor": "cat", "blocks": 1, "tsv_bytes": nbytes,
                              "encode_s": round(t_enc, 3), "decode_s": round(t_dec, 3),
                              "encode_gbs": nbytes / t_enc / 1e9, "decode_gbs": nbytes / t_dec / 1e9}), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
