"""Helpers for the -m gpu parity tests (product through the C ABI vs the oracle)."""
from __future__ import annotations

import oracle as O


def first_diff(a: bytes, b: bytes) -> str:
    n = min(len(a), len(b))
    i = next((k for k in range(n) if a[k] != b[k]), n)
    return (f"len got={len(a)} want={len(b)} first diff at {i}: got={a[max(0, i - 8):i + 24].hex()} "
            f"want={b[max(0, i - 8):i + 24].hex()}")


def oracle_blocks(sch, tsv: bytes, trim=False, rows_per_block=0):
    """Oracle file image split into (header, [block bytes...]) using the decoder's block walk."""
    r = O.encode(sch, tsv, trim=trim, rows_per_block=rows_per_block)
    return r


def split_header(image: bytes):
    _, _, hl = O.read_header(image)
    return image[:hl], image[hl:]


def encode_file_with_product(ctx, sch, tsv: bytes, trim=False, rows_per_block=0, plan=None) -> bytes:
    """File image = oracle-independent header bytes + product blocks (the host stitcher in miniature).
    plan: [(rows, spill_columns), ...] explicit block boundaries; rows left after the plan form one last block."""
    # header: version 11, empty metadata, names, types, char sizes (ConvertToZDW.cpp:673-737)
    hdr = (11).to_bytes(2, "little") + (0).to_bytes(4, "little")
    for nme in sch.names:
        hdr += nme.encode("latin1") + b"\0"
    hdr += b"\0" + bytes(sch.types) + b"".join(int(c).to_bytes(2, "little") for c in sch.charsize)
    out = bytearray(hdr)
    pos = 0
    longest = 0
    plan = list(plan or [])
    k = 0
    while True:
        rows, spill = plan[k] if k < len(plan) else (rows_per_block, 0)
        k += 1
        blk = ctx.encode_block(sch.types, tsv[pos:], trim=trim, prev_longest_line=longest, max_rows=rows, spill_cols=spill)
        if blk.nrows == 0:
            break
        out += blk.data
        longest = blk.longest_line
        pos += blk.tsv_consumed
        if blk.nrows == blk.rows_in_buffer:
            break
    return bytes(out)


_cudart = None


def dev_bytes(ptr: int, n: int) -> bytes:
    """Copies n bytes from a raw device pointer to host (plain cudaMemcpy through libcudart)."""
    import ctypes

    global _cudart
    if _cudart is None:
        import torch  # noqa: F401  (loads the CUDA runtime into the process)
        for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _cudart = ctypes.CDLL(name)
                break
            except OSError:
                continue
        _cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    buf = ctypes.create_string_buffer(max(n, 1))
    rc = _cudart.cudaMemcpy(buf, ctypes.c_void_p(ptr), n, 2)  # cudaMemcpyDeviceToHost
    assert rc == 0, f"cudaMemcpy failed: {rc}"
    return buf.raw[:n]


def decode_file_with_product(ctx, image: bytes, **kw):
    """Walks the blocks of a ZDW file image with the product decoder (host input). Returns (tsv, nblocks, consumed)."""
    sch, ver, hl = O.read_header(image)
    pos = hl
    out = bytearray()
    nblocks = 0
    while True:
        blk = ctx.decode_block(sch.types, image[pos:], at_end_of_file=True, **kw)
        out += blk.tsv
        pos += blk.consumed
        nblocks += 1
        if blk.is_last:
            break
    return bytes(out), nblocks, pos
