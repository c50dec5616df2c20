#!/usr/bin/env python
"""c5_full.py -- config C5 at its full size: 5 x 10^7 rows = 10^8 unique 65-byte strings, about 7.2 GB of
TSV whose dictionary (6.6 GB) does not fit the format's 4 GiB limit, hence >= 2 blocks.

  python tools/c5_full.py [--rows 50000000] [--rows-per-block 25000000] [--ref-decode] [--api]

Rows come from tools/synth_gen.c:c5_rows (the rows of tests/c5_check.py, generated in C on all cores).  Every block is
encoded through the C ABI from host memory, the blocks are stitched into one .zdw file (zdw_b200.shard.stitch_blocks),
decoded again block by block and compared with the source; --ref-decode lets the compiled reference decode OUR file
(streamed compare, about 35 s per 10^7 rows); --api times the row-at-a-time C++ API (zdw_b200/bin/api_rowloop beside
oracle/_ref/api_rowloop).  Prints one JSON line."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--rows-per-block", type=int, default=25_000_000)
    ap.add_argument("--ref-decode", action="store_true")
    ap.add_argument("--api", action="store_true")
    args = ap.parse_args()

    import bench
    import c5_check
    import oracle as O
    from zdw_b200 import Context
    from zdw_b200.shard import stitch_blocks

    bench._build_synth()
    L = C.CDLL(str(ROOT / "tools" / "libsynth_gen.so"))
    L.c5_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]
    L.c5_rows.restype = C.c_size_t
    sch = O.parse_desc(c5_check.DESC)
    res = {"config": "C5 synthetic high-cardinality text, full size", "rows": args.rows, "rows_per_block": args.rows_per_block}

    # ---- generate, one buffer per block, in parallel slices
    t0 = time.time()
    blocks = []
    nthreads = max(1, min(32, os.cpu_count() or 1))
    for first in range(0, args.rows, args.rows_per_block):
        count = min(args.rows_per_block, args.rows - first)
        cap = count * 160 + 4096
        buf = (C.c_uint8 * cap)()
        per = -(-count // nthreads)
        slices = [(first + k * per, min(per, count - k * per)) for k in range(nthreads) if k * per < count]
        tmp = [(C.c_uint8 * (n * 160 + 256))() for _, n in slices]
        with ThreadPoolExecutor(nthreads) as ex:
            lens = list(ex.map(lambda a: L.c5_rows(a[0][0], a[0][1], C.addressof(a[1]), len(a[1])), zip(slices, tmp)))
        at = 0
        for t, n in zip(tmp, lens):
            assert n > 0
            C.memmove(C.addressof(buf) + at, t, n)
            at += n
        del tmp
        blocks.append((buf, at, count))
    res["gen_s"] = round(time.time() - t0, 1)
    res["tsv_bytes"] = sum(b[1] for b in blocks)

    ctx = Context(0)
    # ---- parity anchor: a prefix of the first block against the oracle
    small = bytes(memoryview(blocks[0][0])[:blocks[0][1]][:14_000_000].tobytes().rsplit(b"\n", 1)[0] + b"\n")
    want = O.encode(sch, small)
    _, _, hl = O.read_header(want.data)
    res["prefix_zdw_bit_exact_vs_oracle"] = ctx.encode_block(sch.types, small).data == want.data[hl:]
    file_header = want.data[:hl]

    # ---- encode every block (host memory in, host memory out), stitch
    enc_s, zdw_blocks, info = [], [], []
    tarr = (C.c_uint8 * len(sch.types))(*sch.types)
    from zdw_b200 import capi
    schs = capi._Schema(len(sch.types), C.cast(tarr, C.POINTER(C.c_uint8)))
    for buf, n, count in blocks:
        o = capi._EncOpts(0, 0, 0, 0, 0, 0, 0, 0, 0)
        out = capi._BlockOut()
        t0 = time.time()
        rc = ctx._L.zdwb_encode_block(ctx._h, C.byref(schs), C.c_void_p(C.addressof(buf)), n, C.byref(o), C.byref(out))
        enc_s.append(time.time() - t0)
        if rc:
            raise SystemExit(f"encode failed: {ctx.last_error()}")
        assert out.nrows == count and out.tsv_consumed == n
        zdw_blocks.append(capi._copy_out(out.bytes, out.len))
        info.append({"rows": out.nrows, "dict_entries": out.dict_entries, "dict_bytes": out.dict_bytes, "zdw_bytes": out.len})
    res["blocks"] = info
    res["encode_s_per_block"] = [round(x, 3) for x in enc_s]
    res["encode_gbs_host_to_host"] = round(res["tsv_bytes"] / sum(enc_s) / 1e9, 2)
    image = stitch_blocks(file_header, zdw_blocks)
    del zdw_blocks
    res["zdw_file_bytes"] = len(image)

    # ---- decode block by block through the C ABI, compare with the source
    pos, ok, dec_s = hl, True, []
    for k, (buf, n, count) in enumerate(blocks):
        t0 = time.time()
        d = ctx.decode_block(sch.types, memoryview(image)[pos:pos + info[k]["zdw_bytes"] + 16], at_end_of_file=(k == len(blocks) - 1))
        dec_s.append(time.time() - t0)
        ok = ok and d.nrows == count and d.consumed == info[k]["zdw_bytes"] and d.is_last == (k == len(blocks) - 1)
        ok = ok and len(d.tsv) == n and d.tsv == memoryview(buf)[:n]
        pos += d.consumed
    res["decode_s_per_block"] = [round(x, 3) for x in dec_s]
    res["roundtrip_bit_exact"] = bool(ok and pos == len(image))

    if args.ref_decode and bench.have_ref():
        t0 = time.time()
        res["reference_decodes_our_file_to_the_source"] = bench.ref_decode_matches(image, (memoryview(b[0])[:b[1]] for b in blocks))
        res["reference_decode_s"] = round(time.time() - t0, 1)

    if args.api:
        tmp = tempfile.mkdtemp(dir="/dev/shm" if Path("/dev/shm").is_dir() else None)
        try:
            (Path(tmp) / "c5.zdw").write_bytes(image)
            api = {}
            for name, tool in (("zdw_b200", ROOT / "zdw_b200" / "bin" / "api_rowloop"), ("reference", ROOT / "oracle" / "_ref" / "api_rowloop")):
                if not tool.exists():
                    continue
                pr = subprocess.run([str(tool), "--checksum", "c5.zdw"], cwd=tmp, capture_output=True, text=True, timeout=3000)
                if pr.returncode != 0:
                    api[name] = {"error": pr.stderr[-300:]}
                    continue
                api[name] = json.loads(pr.stdout.strip().splitlines()[-1])
            if "zdw_b200" in api and "reference" in api and "fnv1a" in api["zdw_b200"] and "fnv1a" in api["reference"]:
                api["same_rows"] = (api["zdw_b200"]["rows"], api["zdw_b200"]["fnv1a"]) == (api["reference"]["rows"], api["reference"]["fnv1a"])
            res["unconvert_api"] = api
        finally:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
