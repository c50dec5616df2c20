#!/usr/bin/env python
"""c5_check.py -- config C5: synthetic high-cardinality text, dictionary-bound.

Schema `id int(11) unsigned, k varchar(255), v varchar(255), w varchar(255)`; row i: i+1, 'K'+hex(sha256(i)),
'V'+reversed hex, 'W'+(i mod 97).  Encodes N rows on the GPU as one block (device-resident timing with per-kernel
CUDA events), decodes it again, checks the round trip bit for bit and - for --oracle-rows rows - the .zdw bytes against
the oracle.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

DESC = b"id\tint(11) unsigned\nk\tvarchar(255)\nv\tvarchar(255)\nw\tvarchar(255)\n"


def make_rows(n: int) -> bytes:
    out = []
    for i in range(n):
        h = hashlib.sha256(str(i).encode()).hexdigest()
        out.append(f"{i + 1}\tK{h}\tV{h[::-1]}\tW{i % 97}\n")
    return "".join(out).encode()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--oracle-rows", type=int, default=100_000)
    ap.add_argument("--api", action="store_true", help="also time the getRow loop of the C++ API (ours and the reference's)")
    ap.add_argument("--api-rows", type=int, default=1_000_000)
    ap.add_argument("--api-rows-per-block", type=int, default=0)
    args = ap.parse_args()
    import torch

    import bench
    import oracle as O
    from zdw_b200 import Context

    sch = O.parse_desc(DESC)
    t0 = time.time()
    tsv = make_rows(args.rows)
    gen_s = time.time() - t0
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    res = {"config": "C5 synthetic high-cardinality text", "rows": args.rows, "tsv_bytes": len(tsv), "gen_s": round(gen_s, 1)}

    # parity on a prefix the oracle finishes quickly
    n_or = min(args.oracle_rows, args.rows)
    small = b"".join(tsv.split(b"\n", n_or)[:n_or]) if False else b"\n".join(tsv.split(b"\n")[:n_or]) + b"\n"
    want = O.encode(sch, small)
    _, _, hl = O.read_header(want.data)
    blk = ctx.encode_block(sch.types, small)
    res["oracle_rows"] = n_or
    res["zdw_bit_exact_vs_oracle"] = blk.data == want.data[hl:]
    res["decode_bit_exact_vs_oracle"] = ctx.decode_block(sch.types, blk.data).tsv == O.decode(want.data).tsv == small

    # full size, device resident
    t = torch.empty(len(tsv) + 64, dtype=torch.uint8, device=dev)
    t[:len(tsv)].copy_(torch.frombuffer(bytearray(tsv), dtype=torch.uint8))
    torch.cuda.synchronize()
    ctx.encode_block(sch.types, t.data_ptr(), len(tsv), input_on_device=True, output_on_device=True)  # warm-up
    ctx.set_tuning("kernel_timing", 1)
    ctx.kernel_times()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    blk = ctx.encode_block(sch.types, t.data_ptr(), len(tsv), input_on_device=True, output_on_device=True)
    e1.record()
    e1.synchronize()
    kt = ctx.kernel_times()
    res["encode_ms"] = round(e0.elapsed_time(e1), 3)
    res["encode_gbs"] = round(len(tsv) / e0.elapsed_time(e1) / 1e6, 2)
    res["encode_kernels_ms"] = {k: round(v[1], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])[:8]}
    res["dict_entries"], res["dict_bytes"], res["zdw_bytes"] = blk.dict_entries, blk.dict_bytes, blk.length
    z = torch.empty(blk.length + 64, dtype=torch.uint8, device=dev)
    bench._d2d(torch, z, blk.dev_ptr, blk.length)
    ctx.decode_block(sch.types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)  # warm-up
    ctx.kernel_times()
    e0.record()
    dec = ctx.decode_block(sch.types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)
    e1.record()
    e1.synchronize()
    kt = ctx.kernel_times()
    res["decode_ms"] = round(e0.elapsed_time(e1), 3)
    res["decode_gbs"] = round(len(tsv) / e0.elapsed_time(e1) / 1e6, 2)
    res["decode_kernels_ms"] = {k: round(v[1], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])[:6]}
    back = torch.empty(dec.length, dtype=torch.uint8, device=dev)
    bench._d2d(torch, back, dec.dev_ptr, dec.length)
    res["roundtrip_bit_exact"] = dec.length == len(tsv) and bool(torch.equal(back, t[:len(tsv)]))

    # decode through the row-at-a-time C++ API: the getRow loop of test_unconvert_api without printf,
    # the same source built against this repository's host classes and against the unmodified reference
    if args.api:
        import gpuutil as G
        import shutil
        import subprocess
        import tempfile
        tmp = tempfile.mkdtemp(dir="/dev/shm" if Path("/dev/shm").is_dir() else None)
        try:
            n_api = min(args.api_rows, args.rows)
            part = b"\n".join(tsv.split(b"\n", n_api)[:n_api]) + b"\n"
            image = G.encode_file_with_product(ctx, sch, part, rows_per_block=args.api_rows_per_block)
            (Path(tmp) / "c5.zdw").write_bytes(image)
            api = {"rows": n_api, "zdw_bytes": len(image)}
            for name, tool in (("zdw_b200", ROOT / "zdw_b200" / "bin" / "api_rowloop"), ("reference", ROOT / "oracle" / "_ref" / "api_rowloop")):
                if not tool.exists():
                    api[name] = None
                    continue
                best = None
                for k in range(3):  # run 0 checks the rows (checksum), runs 1-2 time the API alone; best of the two
                    pr = subprocess.run([str(tool)] + (["--checksum"] if k == 0 else []) + ["c5.zdw"], cwd=tmp, capture_output=True,
                                        text=True, timeout=1200)
                    if pr.returncode != 0:
                        raise SystemExit(f"{tool} failed: {pr.stderr[-500:]}")
                    r = json.loads(pr.stdout.strip().splitlines()[-1])
                    if k == 0:
                        fnv = r["fnv1a"]
                    elif best is None or r["mb_per_s"] > best["mb_per_s"]:
                        best = r
                best["fnv1a"] = fnv
                api[name] = best
            if api.get("zdw_b200") and api.get("reference"):
                api["same_rows"] = (api["zdw_b200"]["rows"], api["zdw_b200"]["fnv1a"]) == (api["reference"]["rows"], api["reference"]["fnv1a"])
            res["unconvert_api"] = api
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
