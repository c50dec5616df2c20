"""zdw_b200.desc (the product's host-side .desc.sql parser) against the oracle's restatement of ReadDescFile
(cplusplus/ConvertToZDW.cpp:91-162) on every corpus schema, the three golden sidecars and hand-made oddities."""
import pytest

import corpus
import oracle as O
from zdw_b200 import desc as D

ODD = [
    b"a\tvarchar(70000)\nb\tchar(1)\nc\tchar(2)\nd\tchar(3)\ne\tchar\n",
    b"Field\tType\nfield2\tint(11)\nx\t decimal(24,12)\ny\tdecimal(3,1) unsigned\nz\tXdecimal\n",
    b"a\ttinyint(3) unsigned\nb\ttinyint(4)\nc\tsmallint(5) unsigned\nd\tbigint(20)\ne\tmediumint(9)\nf\tfloat unsigned\n",
    b"a\ttext\nb\ttinytext\nc\tmediumtext\nd\tlongtext\ne\tdatetime\nf\ttimestamp\ng\tdate\nh\tenum('a','b')\n",
    b"a\tvarchar\nb\tvarchar(\nc\tvarchar(-5)\nd\tvarchar( 12)\n",
    b"no_newline_at_end\tint(11)",
    b"",
    b"long" + b"x" * 1100 + b"\tint(11)\nnext\ttext\n",
]


def _same(text):
    try:
        want = O.parse_desc(text)
    except ValueError:
        with pytest.raises(D.DescError):
            D.parse_desc(text)
        return
    got = D.parse_desc(text)
    assert (got.names, got.types, got.charsize) == (want.names, want.types, want.charsize)


@pytest.mark.parametrize("case", corpus.cases(), ids=corpus.case_ids())
def test_corpus_schemas(case):
    _same(case[1])


@pytest.mark.parametrize("name", ["test", "analytics-hits", "movie_tickets"])
def test_golden_sidecars(name):
    _same(O.golden(f"{name}.desc.sql"))


@pytest.mark.parametrize("k", range(len(ODD)))
def test_oddities(k):
    _same(ODD[k])


def test_line_without_tab_is_an_error():
    with pytest.raises(D.DescError):
        D.parse_desc(b"a\tint(11)\nbroken line\n")
    _same(b"a\tint(11)\nbroken line\n")
