"""Pins the oracle port (oracle/zdw_oracle.c) to the reference's own golden vectors
(test-files/test.zdw v9, analytics-hits.zdw v10, movie_tickets.zdw v10; copies under tests/golden/)."""
import pytest

import oracle as O

DATASETS = ["test", "analytics-hits", "movie_tickets"]


@pytest.mark.parametrize("name", DATASETS)
def test_encode_matches_golden(name):
    sch = O.parse_desc(O.golden(f"{name}.desc.sql"))
    tsv = O.golden(f"{name}.sql")
    want = O.golden_to_v11(O.golden(f"{name}.zdw"))
    got = O.encode(sch, tsv)
    assert got.rc == 0
    assert got.nblocks == 1
    assert len(got.data) == len(want)
    assert got.data == want


@pytest.mark.parametrize("name", DATASETS)
def test_decode_golden_roundtrip(name):
    tsv = O.golden(f"{name}.sql")
    z = O.golden(f"{name}.zdw")  # v9 / v10 image as shipped
    got = O.decode(z)
    assert got.rc == 0
    assert got.version in (9, 10)
    assert got.tsv == tsv
    assert got.consumed == len(z)


def test_known_sizes():
    # test-files/README.md: sizes of the TSVs and raw ZDW files
    assert len(O.golden("analytics-hits.sql")) == 14468990
    assert len(O.golden("analytics-hits.zdw")) == 1303440
    assert len(O.golden("movie_tickets.sql")) == 32653800
    assert len(O.golden("movie_tickets.zdw")) == 18290691


def test_worked_example_bytes():
    # SURVEY Appendix A worked example: test.sql -> 139 bytes
    sch = O.parse_desc(O.golden("test.desc.sql"))
    got = O.encode(sch, O.golden("test.sql"))
    assert len(got.data) == 139
    assert got.data[:6] == bytes.fromhex("0b0000000000")
    assert sch.names == ["firstName", "lastName", "age", "eventCode"]
