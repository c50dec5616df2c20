#!/usr/bin/env python
"""profile_one.py -- one device-resident encode + decode of a synthetic C4-shaped block, for use under ncu.

  ncu ... python tools/profile_one.py [--rows N] [--reps K] [--mode both|encode|decode]

Prints the per-kernel CUDA-event times of the last repetition (never a bench value: this runs under a profiler).
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=131072)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--mode", default="both", choices=["both", "encode", "decode"])
    ap.add_argument("--timing", type=int, default=0)
    ap.add_argument("--tune", action="append", default=[], help="name=value tuning knob (repeatable)")
    ap.add_argument("--sweep", default="", help="name=v1,v2,...: time encode + decode kernels for every value of one knob")
    args = ap.parse_args()

    import torch

    import bench
    from zdw_b200 import Context

    synth = bench.Synth()
    types = synth.schema.types
    cap = synth.cap_for(args.rows)
    buf = (C.c_uint8 * cap)()
    n = synth.block_into(0, args.rows, C.addressof(buf), cap)
    dev = torch.device("cuda", 0)
    t = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    t[:n].copy_(torch.frombuffer((C.c_uint8 * n).from_address(C.addressof(buf)), dtype=torch.uint8))
    torch.cuda.synchronize()
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    if args.timing:
        ctx.set_tuning("kernel_timing", 1)
    for kv in args.tune:
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))
    if args.sweep:
        name, vals = args.sweep.split("=")
        ctx.set_tuning("kernel_timing", 1)
        ref = None
        ref_rows = None
        for v in [int(x) for x in vals.split(",")]:
            ctx.set_tuning(name, v)
            best = {}
            for it in range(3):
                blk = ctx.encode_block(types, t.data_ptr(), n, input_on_device=True, output_on_device=True)
                for k, (cnt, ms) in ctx.kernel_times().items():
                    best[k] = min(best.get(k, 1e9), ms)
                z = torch.empty(blk.length + 64, dtype=torch.uint8, device=dev)
                bench._d2d(torch, z, blk.dev_ptr, blk.length)
                if ref is None:
                    ref = z[:blk.length].clone()
                else:
                    assert ref.numel() == blk.length and bool((ref == z[:blk.length]).all()), "encoded block differs between knob values"
                dec = ctx.decode_block(types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)
                for k, (cnt, ms) in ctx.kernel_times().items():
                    best[k] = min(best.get(k, 1e9), ms)
                if it == 0:  # the decoded rows must not depend on the knob either
                    rows_out = torch.empty(dec.length, dtype=torch.uint8, device=dev)
                    bench._d2d(torch, rows_out, dec.dev_ptr, dec.length)
                    if ref_rows is None:
                        ref_rows = rows_out
                    else:
                        assert ref_rows.numel() == dec.length and bool((ref_rows == rows_out).all()), "decoded rows differ between knob values"
            enc = sum(ms for k, ms in best.items() if not k.startswith(("k_dec", "k_carry", "k_dict_nulmap")))
            top = sorted(best.items(), key=lambda kv: -kv[1])[:6]
            print(f"{name}={v}: sum_ms={sum(best.values()):.3f} " + " ".join(f"{k}={ms:.3f}" for k, ms in top), flush=True)
        return
    blk = ctx.encode_block(types, t.data_ptr(), n, input_on_device=True, output_on_device=True)
    z = torch.empty(blk.length + 64, dtype=torch.uint8, device=dev)
    bench._d2d(torch, z, blk.dev_ptr, blk.length)
    for _ in range(args.reps):
        if args.mode in ("both", "encode"):
            ctx.encode_block(types, t.data_ptr(), n, input_on_device=True, output_on_device=True)
            if args.timing:
                print("encode", sorted(ctx.kernel_times().items(), key=lambda kv: -kv[1][1]))
        if args.mode in ("both", "decode"):
            ctx.decode_block(types, z.data_ptr(), blk.length, input_on_device=True, output_on_device=True)
            if args.timing:
                print("decode", sorted(ctx.kernel_times().items(), key=lambda kv: -kv[1][1]))
    torch.cuda.synchronize()
    print(f"rows={args.rows} tsv={n} zdw={blk.length} launches={ctx.kernel_launches()}")


if __name__ == "__main__":
    main()
