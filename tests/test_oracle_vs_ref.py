"""Differential test: oracle port vs the compiled, unmodified reference (oracle/_ref) on the
adversarial corpus.  Skipped when oracle/_ref was not built (it is built wherever /root/reference
is mounted and travels to the GPU box with the snapshot)."""
import pytest

import corpus
import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")

FAST = [c for c in corpus.cases() if not c[0].startswith("d2_")]
D2 = [c for c in corpus.cases() if c[0].startswith("d2_")]


def _check(case):
    name, desc, tsv, opts = case
    args = ["-q"] + (["-t"] if opts.get("trim") else [])
    rc, zref, log = O.ref_encode(tsv, desc, args)
    sch = O.parse_desc(desc)
    got = O.encode(sch, tsv, trim=bool(opts.get("trim")))
    if rc != 0:
        assert got.rc == 15, (name, rc, log)
        assert f"Row {got.bad_row} had the problem" in log
        if "expect_rc" in opts:
            assert opts["expect_rc"] == 15
        return
    assert "expect_rc" not in opts, name
    assert got.rc == 0, name
    assert zref is not None
    assert got.data == zref, name
    if got.total_rows == 0:
        return  # header-only file: the reference decoder throws on it (SURVEY B-15)
    rc2, tsv_ref, err = O.ref_decode(zref)
    dec = O.decode(zref)
    if rc2 != 0:
        # e.g. a block whose columns are all unused has zero-byte rows the reference cannot read (ROW_COUNT_ERR)
        assert dec.rc == rc2, (name, err)
        return
    assert dec.rc == 0
    assert dec.tsv == tsv_ref, name


@pytest.mark.parametrize("case", FAST, ids=[c[0] for c in FAST])
def test_corpus_case(case):
    _check(case)


def test_corpus_tile_boundaries():
    # d2: several hundred tiny cases, checked against one oracle run each (sampled to keep CPU time low)
    for case in D2[::7]:
        _check(case)


def test_goldens_through_ref():
    for name in ("test", "analytics-hits"):
        desc = O.golden(f"{name}.desc.sql")
        tsv = O.golden(f"{name}.sql")
        rc, zref, log = O.ref_encode(tsv, desc, ["-q"])
        assert rc == 0, log
        assert zref == O.golden_to_v11(O.golden(f"{name}.zdw"))
        rc2, back, _ = O.ref_decode(zref)
        assert rc2 == 0 and back == tsv


def test_column_selection_vs_ref():
    desc = O.golden("test.desc.sql")
    tsv = O.golden("test.sql")
    z = O.encode(O.parse_desc(desc), tsv).data
    # -c eventCode,firstName : reorder (SURVEY appendix B probe list)
    rc, out, _ = O.ref_decode(z, ["-c", "eventCode,firstName"])
    assert rc == 0
    got = O.decode(z, out_col=[1, -1, -1, 0], n_out=2)
    assert got.tsv == out


def test_api_decode_vs_ref():
    for name in ("test", "analytics-hits"):
        z = O.golden_to_v11(O.golden(f"{name}.zdw"))
        rc, out, _ = O.ref_api_decode(z)
        assert rc == 0
        assert out == O.golden(f"{name}.sql")


@pytest.mark.parametrize("case", corpus.big_cases(), ids=[c[0] for c in corpus.big_cases()])
def test_restatement_equals_reference_on_big_cases(case):
    """App. D d9 at the 3- / 4-byte dictionary offset edge (2^24 dictionary bytes)."""
    name, desc, tsv, _ = case
    sch = O.parse_desc(desc)
    want = O.encode(sch, tsv)
    rc, ref_zdw, log = O.ref_encode(tsv, desc, ["-q"], timeout=300)
    assert rc == 0 and want.rc == 0, log
    assert want.data == ref_zdw, name
    expect_idx = 3 if name.endswith("16777215") else 4
    _, _, hl = O.read_header(want.data)
    assert want.data[hl + 9] == expect_idx
    assert O.decode(want.data).tsv == tsv


def test_longest_line_is_cumulative_over_blocks():
    desc, tsv, rpb = corpus.later_block_long_line()
    sch = O.parse_desc(desc)
    img = O.encode(sch, tsv, rows_per_block=rpb).data
    dec = O.decode(img)
    assert dec.nblocks == 4 and dec.tsv == tsv
    import struct
    lines = [struct.unpack_from("<I", img, off + 4)[0] for off in dec.block_offset]
    # (a block cut by a row count has read the row behind it already - its length counts: 32768 for block 0)
    assert lines == [32768, 32768, 131072, 131072]
