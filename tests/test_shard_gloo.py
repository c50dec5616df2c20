"""N>1 path on CPU: two gloo ranks each encode their contiguous block range (the oracle stands in for the GPU
encoder - this test is about the sharding / stitching host logic), rank 0 gathers the blocks, stitches them and must
reproduce the single-process multi-block file bit for bit."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _split_rows(tsv: bytes, rows_per_block: int):
    lines = tsv.split(b"\n")[:-1]
    return [b"\n".join(lines[i:i + rows_per_block]) + b"\n" for i in range(0, len(lines), rows_per_block)]


def _worker(rank, world, port, rows_per_block, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import oracle as O
    from zdw_b200.shard import block_range, stitch_blocks

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        desc, tsv = O.golden("movie_tickets.desc.sql"), O.golden("movie_tickets.sql")[:3_000_000].rsplit(b"\n", 1)[0] + b"\n"
        sch = O.parse_desc(desc)
        shards = _split_rows(tsv, rows_per_block)
        mine = []
        for b in block_range(rank, world, len(shards)):
            enc = O.encode(sch, shards[b])  # one block, as a rank's GPU would produce it (prev_longest_line = 0)
            _, _, hl = O.read_header(enc.data)
            mine.append((b, enc.data[hl:], enc.data[:hl]))
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank == 0:
            allb = sorted(x for part in gathered for x in part)
            assert [x[0] for x in allb] == list(range(len(shards)))
            image = stitch_blocks(allb[0][2], [x[1] for x in allb])
            want = O.encode(sch, tsv, rows_per_block=rows_per_block)
            q.put((image == want.data, want.nblocks, len(shards), O.decode(image).tsv == tsv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows_per_block", [(2, 10000), (2, 2500), (3, 5000)])
def test_two_rank_block_sharding_and_stitching(world, rows_per_block):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows_per_block, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same, nblocks, nshards, roundtrip = q.get(timeout=10)
    assert same and roundtrip and nblocks == nshards


def test_block_range_partitions_every_block_once():
    from zdw_b200.shard import block_range
    for world in (1, 2, 3, 4, 8):
        for nb in (1, 7, 8, 64, 65):
            seen = [b for r in range(world) for b in block_range(r, world, nb)]
            assert seen == list(range(nb))


def test_stitch_patches_is_last_and_cumulative_longest_line():
    import struct
    from zdw_b200.shard import stitch_blocks
    mk = lambda rows, line, last: struct.pack("<II", rows, line) + bytes([last]) + b"\x00" + b"\x00" * 3
    img = stitch_blocks(b"HDR", [mk(5, 32768, 1), mk(6, 16384, 1), mk(7, 65536, 0)])
    blocks = [img[3 + i * 13: 3 + (i + 1) * 13] for i in range(3)]
    assert [struct.unpack_from("<II", b) for b in blocks] == [(5, 32768), (6, 32768), (7, 65536)]
    assert [b[8] for b in blocks] == [0, 0, 1]
