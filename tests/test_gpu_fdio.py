"""The host<->device paths that do not start from pinned caller memory (ABI 4): pageable buffers of more than 32 MiB go
through the context's pinned ring inside zdwb_encode_block / zdwb_decode_block, and zdwb_fd_to_device /
zdwb_device_to_fd move a window of a file to the device and decoded rows to their place in a file (what the host tools'
workers use).  Every result is compared with the plain path on the same C4-shaped rows (which tests/test_gpu_c4.py pins
to the reference)."""
import ctypes as C
import os

import pytest

import gpuutil as G
import oracle as O

pytestmark = pytest.mark.gpu

ROWS = 12000  # ~46 MB of TSV: above the 32 MiB threshold of the ring paths, several 8 MiB chunks


@pytest.fixture(scope="module")
def synth():
    import bench
    return bench.Synth()


@pytest.fixture(scope="module")
def case(synth):
    cap = synth.cap_for(ROWS)
    buf = (C.c_uint8 * cap)()
    n = synth.block_into(5, ROWS, C.addressof(buf), cap)
    tsv = bytes(memoryview(buf)[:n])
    assert len(tsv) > (40 << 20)
    sch = O.parse_desc(synth.desc)
    port = O.encode(sch, tsv)
    assert port.rc == 0 and port.total_rows == ROWS
    _, _, hl = O.read_header(port.data)
    return sch, tsv, port.data[hl:]  # the block without the file header


def test_pageable_buffers_take_the_ring(case):
    from zdw_b200 import Context
    sch, tsv, want_block = case
    with Context(0) as ctx:
        blk = ctx.encode_block(sch.types, tsv)  # bytes = pageable memory, > 32 MiB: ring_h2d
        assert blk.data == want_block, G.first_diff(blk.data, want_block)
        dec = ctx.decode_block(sch.types, blk.data)  # first calls of a context return a plain buffer: ring_d2h
        assert dec.tsv == tsv, G.first_diff(dec.tsv, tsv)
        for _ in range(3):  # later calls: the pinned result buffer (the usual copy), same bytes
            dec = ctx.decode_block(sch.types, blk.data)
        assert dec.tsv == tsv


def test_fd_to_device_and_back(case, tmp_path):
    from zdw_b200 import Context
    sch, tsv, want_block = case
    src = tmp_path / "in.sql"
    junk = b"#" * 4099  # the window does not start at the front of the file, nor at an aligned offset
    src.write_bytes(junk + tsv + b"tail that is not part of the window\n")
    dst = tmp_path / "out.sql"
    with Context(0) as ctx:
        fd = os.open(src, os.O_RDONLY)
        try:
            dev = ctx.fd_to_device(fd, len(junk), len(tsv))
            blk = ctx.encode_block(sch.types, dev, len(tsv), input_on_device=True)
            assert blk.data == want_block, G.first_diff(blk.data, want_block)
            # a second window reuses the device buffer
            half = tsv[:tsv.rfind(b"\n", 0, len(tsv) // 2) + 1]
            dev2 = ctx.fd_to_device(fd, len(junk), len(half))
            assert dev2 == dev
            blk2 = ctx.encode_block(sch.types, dev2, len(half), input_on_device=True)
            assert blk2.nrows == half.count(b"\n")
            # the file is shorter than what is asked for
            from zdw_b200.capi import ZdwError
            with pytest.raises(ZdwError) as e:
                ctx.fd_to_device(fd, len(junk), len(tsv) + (1 << 20))
            assert "ended early" in str(e.value)
        finally:
            os.close(fd)
        dec = ctx.decode_block(sch.types, blk.data, output_on_device=True)
        out = os.open(dst, os.O_CREAT | os.O_RDWR, 0o644)
        try:
            os.pwrite(out, b"HEAD", 0)
            ctx.device_to_fd(dec.dev_ptr, dec.length, out, 4)
        finally:
            os.close(out)
        got = dst.read_bytes()
        assert got[:4] == b"HEAD" and got[4:] == tsv, G.first_diff(got[4:], tsv)
        # a pipe: offset < 0 = plain write()
        r, w = os.pipe()
        import threading
        chunks = []

        def drain():
            while True:
                b = os.read(r, 1 << 20)
                if not b:
                    break
                chunks.append(b)
        t = threading.Thread(target=drain)
        t.start()
        try:
            ctx.device_to_fd(dec.dev_ptr, dec.length, w, -1)
        finally:
            os.close(w)
            t.join()
            os.close(r)
        assert b"".join(chunks) == tsv
