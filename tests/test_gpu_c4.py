"""Config C4 - the configuration every bench.py number is quoted on - against the reference, bit for bit.

C4 blocks come from the same generator bench.py uses (`bench.Synth`, seed 20190901).  The CUDA encoder's `.zdw` must equal
what the unmodified compiled reference (`oracle/_ref/convertDWfile`) and the C restatement write for the same rows, and
the reference's `unconvertDWfile` must turn OUR file back into the source rows (SURVEY 8(d) "parity gate for every
timing").  One block of the full bench size (131 072 rows, 0.5 GB) is part of the matrix; the multi-block cases cover
what a multi-GPU run produces: blocks encoded by different contexts, stitched in file order (`isLast`, cumulative
`longestLine`: ConvertToZDW.cpp:841-842,965).
"""
import ctypes as C

import pytest

import gpuutil as G
import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    import bench
    return bench.Synth()


@pytest.fixture(scope="module")
def ctx():
    from zdw_b200 import Context
    c = Context(0)
    yield c
    c.close()


def c4_block(synth, block: int, rows: int) -> bytes:
    cap = synth.cap_for(rows)
    buf = (C.c_uint8 * cap)()
    n = synth.block_into(block, rows, C.addressof(buf), cap)
    return bytes(memoryview(buf)[:n])


@pytest.mark.parametrize("rows,block", [(4096, 0), (32768, 3), (131072, 0)])
def test_c4_block_bit_exact_vs_reference(ctx, synth, rows, block):
    tsv = c4_block(synth, block, rows)
    sch = O.parse_desc(synth.desc)
    ours = G.encode_file_with_product(ctx, sch, tsv)
    port = O.encode(sch, tsv)
    assert port.rc == 0 and port.total_rows == rows
    assert ours == port.data, "vs C restatement: " + G.first_diff(ours, port.data)
    back, nblocks, consumed = G.decode_file_with_product(ctx, ours)
    assert nblocks == 1 and consumed == len(ours)
    assert back == tsv, "our decode of our file: " + G.first_diff(back, tsv)
    if O.have_ref():
        rc, ref_zdw, log = O.ref_encode(tsv, synth.desc, ["-q"], timeout=600)
        assert rc == 0, log
        assert ours == ref_zdw, "vs compiled reference: " + G.first_diff(ours, ref_zdw)
        rc, ref_tsv, err = O.ref_decode(ours, timeout=600)
        assert rc == 0, err
        assert ref_tsv == tsv, "reference decode of our file: " + G.first_diff(ref_tsv, tsv)


def test_c4_blocks_from_two_contexts_stitched(synth):
    """Six 8 192-row blocks, encoded alternately by two contexts the way two ranks would (each with
    prev_longest_line = 0), stitched on the host: the file equals the single-process multi-block file of the
    restatement, and the reference decodes it to the source rows."""
    from zdw_b200 import Context
    from zdw_b200.shard import block_range, stitch_blocks
    rows, nb = 8192, 6
    sch = O.parse_desc(synth.desc)
    parts = [c4_block(synth, 10 + b, rows) for b in range(nb)]
    want = O.encode(sch, b"".join(parts), rows_per_block=rows)
    assert want.rc == 0 and want.nblocks == nb
    _, _, hl = O.read_header(want.data)
    ctxs = [Context(0), Context(0)]
    try:
        blocks = [None] * nb
        for rank, c in enumerate(ctxs):
            for b in block_range(rank, len(ctxs), nb):
                blocks[b] = c.encode_block(sch.types, parts[b]).data
        image = stitch_blocks(want.data[:hl], blocks)
        assert image == want.data, G.first_diff(image, want.data)
        back, nblocks, consumed = G.decode_file_with_product(ctxs[1], image)
        assert nblocks == nb and consumed == len(image) and back == b"".join(parts)
    finally:
        for c in ctxs:
            c.close()
    if O.have_ref():
        rc, ref_tsv, err = O.ref_decode(image, timeout=600)
        assert rc == 0, err
        assert ref_tsv == b"".join(parts)
