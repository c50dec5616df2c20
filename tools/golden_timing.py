#!/usr/bin/env python
"""golden_timing.py -- device-resident encode / decode times of the reference's own fixtures (configs C2 movie_tickets,
C3 analytics-hits) with per-kernel CUDA-event times.  Prints one JSON line per fixture.  Arguments: name=value tuning knobs."""
from __future__ import annotations

import json
import lzma
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from zdw_b200 import Context
    from zdw_b200.desc import parse_desc

    dev = torch.device("cuda", 0)
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    tune = {}
    for kv in sys.argv[1:]:  # name=value tuning knobs
        k, v = kv.split("=")
        if k != "all":  # all=1: every kernel in the list, not the top six
            ctx.set_tuning(k, int(v))
        tune[k] = int(v)
    for name in ("movie_tickets", "analytics-hits"):
        g = ROOT / "tests" / "golden"
        tsv = lzma.decompress((g / f"{name}.sql.xz").read_bytes())
        sch = parse_desc((g / f"{name}.desc.sql").read_bytes())
        t = torch.empty(len(tsv) + 64, dtype=torch.uint8, device=dev)
        t[:len(tsv)].copy_(torch.frombuffer(bytearray(tsv), dtype=torch.uint8))
        torch.cuda.synchronize()
        res = {"fixture": name, "tsv_bytes": len(tsv)}
        if tune:
            res["tune"] = tune
        for _ in range(2):
            blk = ctx.encode_block(sch.types, t.data_ptr(), len(tsv), input_on_device=True, output_on_device=True)
        z = torch.empty(blk.length + 64, dtype=torch.uint8, device=dev)
        import bench
        bench._d2d(torch, z, blk.dev_ptr, blk.length)
        ctx.decode_block(sch.types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)
        best = {}
        for mode in ("encode", "decode"):
            ms = []
            for k in range(5):
                if k == 4:
                    ctx.set_tuning("kernel_timing", 1)
                    ctx.kernel_times()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if mode == "encode":
                    ctx.encode_block(sch.types, t.data_ptr(), len(tsv), input_on_device=True, output_on_device=True)
                else:
                    ctx.decode_block(sch.types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)
                e1.record()
                e1.synchronize()
                if k < 4:
                    ms.append(e0.elapsed_time(e1))
            kt = ctx.kernel_times()
            ctx.set_tuning("kernel_timing", 0)
            res[mode] = {"ms": round(min(ms), 3), "gbs": round(len(tsv) / min(ms) / 1e6, 2),
                         "kernels_sum_ms": round(sum(v[1] for v in kt.values()), 3), "launches": sum(v[0] for v in kt.values()),
                         "kernels_ms": {k: round(v[1], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])[:(40 if tune.get("all") else 6)]}}
        res["zdw_bytes"] = blk.length
        print(json.dumps(res))


if __name__ == "__main__":
    main()
