"""The C restatement against the compiled reference on C4-shaped rows (the benchmarked config): pins the oracle on the
data the GPU parity tests of tests/test_gpu_c4.py and bench.py's parity gate compare against."""
import ctypes as C

import pytest

import oracle as O


def _c4(rows, block):
    import bench
    synth = bench.Synth()
    cap = synth.cap_for(rows)
    buf = (C.c_uint8 * cap)()
    n = synth.block_into(block, rows, C.addressof(buf), cap)
    return synth, bytes(memoryview(buf)[:n])


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not available")
@pytest.mark.parametrize("rows,block", [(2048, 0), (3000, 63)])
def test_restatement_equals_reference_on_c4_rows(rows, block):
    synth, tsv = _c4(rows, block)
    sch = O.parse_desc(synth.desc)
    port = O.encode(sch, tsv)
    rc, ref_zdw, log = O.ref_encode(tsv, synth.desc, ["-q"])
    assert rc == 0, log
    assert port.rc == 0 and port.data == ref_zdw
    rc, ref_tsv, err = O.ref_decode(port.data)
    assert rc == 0 and ref_tsv == tsv == O.decode(port.data).tsv


def test_c4_generator_is_deterministic_and_canonical():
    """Same (seed, block, rows) -> same bytes; rows survive the oracle round trip unchanged (canonical numerics)."""
    synth, a = _c4(512, 5)
    _, b = _c4(512, 5)
    assert a == b and a.count(b"\n") == 512
    sch = O.parse_desc(synth.desc)
    assert O.decode(O.encode(sch, a).data).tsv == a
