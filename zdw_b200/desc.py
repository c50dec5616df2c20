"""`.desc.sql` schema sidecar -> column names, ZDW type ids, char sizes (host side, Python mirror of
`ConvertToZDW::readDescFile` in zdw_b200/host/ConvertToZDW.cpp; rules of the reference: cplusplus/ConvertToZDW.cpp:91-162,
SURVEY App. B-1).

Prefix matching, in this order: varchar(N); char(1) -> CHAR, char(2) -> CHAR_2, any other char(N) -> VARCHAR; text;
tinytext; mediumtext; longtext; datetime; decimal (also one character in); everything else is an integer, signed unless
the line mentions "unsigned": tinyint / smallint / bigint, and LONG for the rest (int, mediumint, float, double,
timestamp, ...).  Lines starting with "Field" (any case) are skipped; a line without a tab is an error; lines are cut at
1023 bytes like the reference's fgets buffer.
"""
from __future__ import annotations

from dataclasses import dataclass, field

# type ids stored on disk (cplusplus/zdw/zdw_column_type_constants.h:17-35; include/zdw_b200.h)
VARCHAR, TEXT, DATETIME, CHAR_2, VISID_LOW, VISID_HIGH, CHAR, TINY, SHORT, LONG, LONGLONG, DECIMAL = range(12)
TINY_SIGNED, SHORT_SIGNED, LONG_SIGNED, LONGLONG_SIGNED, TINYTEXT, MEDIUMTEXT, LONGTEXT = range(12, 19)


class DescError(ValueError):
    """DESC_FILE_MISSING_TYPE_INFO: a line without a tab."""


@dataclass
class Schema:
    names: list = field(default_factory=list)
    types: list = field(default_factory=list)
    charsize: list = field(default_factory=list)

    @property
    def ncols(self) -> int:
        return len(self.types)


def _atoi(b: bytes) -> int:
    """C atoi: optional white space, optional sign, digits; 0 when there are none."""
    i, n = 0, len(b)
    while i < n and b[i] in b" \t\n\v\f\r":
        i += 1
    neg = False
    if i < n and b[i] in b"+-":
        neg = b[i] == 0x2D
        i += 1
    v = 0
    while i < n and 0x30 <= b[i] <= 0x39:
        v = v * 10 + (b[i] - 0x30)
        i += 1
    return -v if neg else v


def _fgets_lines(text: bytes, cap: int = 1024):
    """The byte strings successive fgets(line, cap, f) calls return."""
    pos, n = 0, len(text)
    while pos < n:
        nl = text.find(b"\n", pos, pos + cap - 1)
        end = nl + 1 if nl >= 0 else min(n, pos + cap - 1)
        yield text[pos:end]
        pos = end


def parse_desc(text: bytes) -> Schema:
    out = Schema()
    for line in _fgets_lines(text):
        if line[:5].lower() == b"field":
            continue
        tab = line.find(b"\t")
        if tab < 0:
            raise DescError("desc line without a tab")
        nul = line.find(b"\0")  # the reference works on C strings
        name = line[:tab] if nul < 0 or nul > tab else line[:nul]
        typ = line[tab + 1:]
        if nul > tab:
            typ = line[tab + 1:nul]
        size = 0
        if typ.startswith(b"varchar"):
            tid, size = VARCHAR, _atoi(typ[8:])
        elif typ.startswith(b"char"):
            size = _atoi(typ[5:])
            tid = CHAR if size == 1 else CHAR_2 if size == 2 else VARCHAR
        elif typ.startswith(b"text"):
            tid = TEXT
        elif typ.startswith(b"tinytext"):
            tid = TINYTEXT
        elif typ.startswith(b"mediumtext"):
            tid = MEDIUMTEXT
        elif typ.startswith(b"longtext"):
            tid = LONGTEXT
        elif typ.startswith(b"datetime"):
            tid = DATETIME
        elif typ.startswith(b"decimal") or typ[1:].startswith(b"decimal"):
            tid = DECIMAL
        else:
            signed = b"unsigned" not in typ
            if typ.startswith(b"tinyint"):
                tid = TINY_SIGNED if signed else TINY
            elif typ.startswith(b"smallint"):
                tid = SHORT_SIGNED if signed else SHORT
            elif typ.startswith(b"bigint"):
                tid = LONGLONG_SIGNED if signed else LONGLONG
            else:
                tid = LONG_SIGNED if signed else LONG
        out.names.append(name.decode("latin1"))
        out.types.append(tid)
        out.charsize.append(size & 0xFFFF)
    return out
