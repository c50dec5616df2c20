"""Block sharding across ranks and in-order stitching (multi-GPU path).

ZDW blocks are self-contained - own dictionary, baselines, previous-row state reset (reference
ConvertToZDW.cpp:504-505,880; decoder UnconvertFromZDW.cpp:985-986) - so whole blocks are the shard unit and no
data-path collective exists: rank r encodes a contiguous range of blocks, the blocks are concatenated in file order
and two header fields are patched per block: `isLast` (1 only on the final block, ConvertToZDW.cpp:841-842) and
`longestLine` (cumulative over the file because m_LongestLine is never reset, ConvertToZDW.cpp:965,
getnextrow.cpp:57-65).
"""
from __future__ import annotations

import struct


def block_range(rank: int, world: int, nblocks: int) -> range:
    """Contiguous, near-equal block ranges in file order; every block belongs to exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank * nblocks // world, (rank + 1) * nblocks // world)


def stitch_blocks(file_header: bytes, blocks: list) -> bytes:
    """file header + blocks (each as returned by zdwb_encode_block with prev_longest_line = 0) -> .zdw image."""
    out = [file_header]
    longest = 0
    for k, b in enumerate(blocks):
        if len(b) < 9:
            raise ValueError("block too short")
        nrows, line = struct.unpack_from("<II", b, 0)
        longest = max(longest, line)
        out.append(struct.pack("<II", nrows, longest) + (b"\x01" if k == len(blocks) - 1 else b"\x00") + bytes(b[9:]))
    return b"".join(out)
