"""Adversarial TSV corpus (cases d1-d15).  Each case is tiny and deterministic;
the same cases are diffed (a) oracle port vs compiled reference on CPU and (b) CUDA product vs
oracle on the GPU.  Cases are (name, desc_bytes, tsv_bytes, opts) with opts a dict of
{trim: bool, expect_rc: int}.
"""
from __future__ import annotations

import random

BS = b"\\"


def desc(cols):
    """cols: list of (name, sqltype) -> .desc.sql bytes."""
    return b"".join(f"{n}\t{t}\n".encode() for n, t in cols)


def rows(rs):
    return b"".join(b"\t".join(r) + b"\n" for r in rs)


def _cases():
    out = []

    def add(name, d, t, **opts):
        out.append((name, d, t, opts))

    two_text = desc([("a", "varchar(255)"), ("b", "text")])

    # d1: escaped tab / newline / backslash runs of length 1..5 before \t and \n, at field start/end, row start
    rs = []
    for k in range(0, 6):
        run = BS * k
        rs.append(b"x" + run + b"\tq" + run + b"\n")          # run before separator / terminator
        rs.append(run + b"\ty" + b"\t" + run + b"z\n")         # run at row start / field start
        rs.append(b"m" + run + b"\nn\tend" + run + b"\n")      # newline after run inside first field
    t = b"".join(rs)
    add("d1_escape_runs", two_text, t)  # some rows will have wrong counts -> oracle/ref agree on rc
    good = []
    for k in (1, 3, 5):
        run = BS * k
        good.append(b"he" + run + b"\tllo\tw" + run + b"\norld\n")
        good.append(run + b"\tlead\t" + run + b"\ntrail" + BS * (k + 1) + b"\n")
    for k in (0, 2, 4):
        run = BS * k
        good.append(b"ev" + run + b"\ten" + run + b"\n")
    add("d1_escape_runs_valid", two_text, b"".join(good))

    # d2: backslash runs crossing power-of-two offsets (GPU tile boundaries)
    for tile in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
        for delta in (-2, -1, 0, 1, 2):
            for runlen in (1, 2, 3, 33, 64, 65):
                pad = tile + delta - runlen - 1
                if pad < 1:
                    continue
                body = b"p" * pad + BS * runlen
                # odd run escapes the tab -> need another real tab
                t = body + b"\tmid\tfin\n" if runlen % 2 else body + b"\tfin\n"
                t += b"second\trow\n"
                add(f"d2_tile{tile}_d{delta}_r{runlen}", two_text, t)

    # d3: blank lines, lone-char last line, missing final newline, CRLF
    add("d3_blank_lines", desc([("a", "int(11)"), ("b", "int(11)")]), b"5\t0\n\n7\t0\n\n\n9\t0\n11\t0")
    add("d3_lone_char_tail", two_text, b"a\tb\nc\td\nx")
    add("d3_crlf", two_text, b"a\tb\r\nc\td\r\n")
    add("d3_leading_blank", two_text, b"\n\n\na\tb\nc\td\n\n")
    add("d3_escaped_nl_then_blank", two_text, b"a" + BS + b"\n\n\tb\nc\td\n")
    add("d3_tail_escaped_nl", two_text, b"a\tb\nc\td" + BS + b"\n")

    # d4: long lines around the row-buffer doubling points
    for L in (16382, 16383, 16384, 32767, 32768, 70000):
        t = b"s\tt\n" + b"x" * (L - 3) + b"\ty\n" + b"u\tv\n"   # middle line has exactly L bytes incl. newline
        add(f"d4_len{L}", two_text, t)
    add("d4_long_unterminated_tail", two_text, b"s\tt\n" + b"x" * 16383 + b"\ty")
    add("d4_long_tail_exact", two_text, b"s\tt\n" + b"x" * 16380 + b"\ty")

    # d5: integer parsing quirks in every numeric type
    ints = [b"", b"0", b"007", b"+4", b"-5", b"  12", b"12abc", b"3.9", b"18446744073709551615",
            b"18446744073709551616", b"-18446744073709551616", b"-0", b"+", b"-", b"0x10", b"1e3", b" -7 ",
            b"\x0b13", b"9223372036854775807", b"9223372036854775808", b"-9223372036854775808", b"255", b"256",
            b"65535", b"65536", b"99999999999", b"1", b"+-3", b"--3", b" +9", b"00000000000000000000001",
            b"184467440737095516150", b"\x0c\r 42"]
    num_types = ["tinyint(3) unsigned", "tinyint(4)", "smallint(5) unsigned", "smallint(6)", "int(11) unsigned",
                 "int(11)", "bigint(20) unsigned", "bigint(20)", "float", "double", "timestamp", "mediumint(8)"]
    d = desc([(f"n{i}", ty) for i, ty in enumerate(num_types)])
    rng = random.Random(5)
    rs = [[v] * len(num_types) for v in ints]
    for _ in range(40):
        rs.append([rng.choice(ints) for _ in num_types])
    add("d5_ints", d, rows(rs))
    add("d5_int_trailing_cr", desc([("a", "int(11)"), ("b", "int(11)")]), b"1\t2\r\n3\t4\r\n")

    # d6: char(1)/char(2)/char(3)
    d = desc([("c1", "char(1)"), ("c2", "char(2)"), ("c3", "char(3)"), ("k", "int(11)")])
    vals = [b"", b"x", b"yz", BS + b"\t", BS + BS, b"\xe9", "é".encode(), BS + b"n", BS, b"0", b" ", BS + b"\n",
            b"\x7f", b"\x80", b"\xff", BS + b"\xff", b"ab", b"~"]
    rs = [[v, v, v, b"%d" % i] for i, v in enumerate(vals)]
    rs += [[rng.choice(vals), rng.choice(vals), rng.choice(vals), b"%d" % rng.randrange(5)] for _ in range(40)]
    add("d6_chars", d, rows(rs))
    add("d6_char_only_neg", desc([("c1", "char(1)")]), b"\xe9\n\xe9\n\xf0\n")

    # d7: decimal / datetime / shared strings across text-like columns
    d = desc([("dec", "decimal(24,12)"), ("dt", "datetime"), ("vc", "varchar(64)"), ("tx", "text"),
              ("tt", "tinytext"), ("mt", "mediumtext"), ("lt", "longtext"), ("dec2", " decimal(10,2)")])
    pool = [b"", b"0.000000000000", b"1.50", b"-2", b"2019-09-01 00:00:00", b"shared", b"2", b"3", b"4", b"ab", b"zz",
            "é".encode(), "ünï".encode(), b"hello", b"he" + BS + b"\tllo", b"line" + BS + b"\nbreak",
            b"trail" + BS + BS, b"q", b"A", b"a", b"aa", b"a\x01", b"a\xff", b"abcdefgh", b"abcdefghi", b"abcdefg",
            b"abcdefgh\x01", b"http://www.example.com/a", b"http://www.example.com/b", b"http://www.example.com/"]
    rs = [[rng.choice(pool) for _ in range(8)] for _ in range(120)]
    add("d7_text_like", d, rows(rs))

    # d8: never populated / only zeros / flag-byte edges 1, 8, 9, 256, 257 used columns
    d = desc([("e", "varchar(10)"), ("z", "int(11)"), ("v", "int(11)"), ("e2", "text"), ("c", "char(1)")])
    add("d8_unused_cols", d, rows([[b"", b"0", b"%d" % (i % 3 + 1), b"", b""] for i in range(10)]))
    add("d8_all_unused", d, rows([[b"", b"0", b"", b"", b""] for _ in range(5)]))
    for nu in (1, 7, 8, 9, 16, 17, 255, 256, 257):
        cols = [(f"c{i}", "int(11)" if i % 2 else "varchar(20)") for i in range(nu)] + [("pad", "text")]
        rs = []
        for r in range(9):
            rs.append([(b"%d" % (1 + (r * 7 + i) % 5)) if (r + i) % 3 else (b"%d" % (i + 1)) for i in range(nu)] + [b""])
        add(f"d8_used{nu}", desc(cols), rows(rs))

    # d9: dictionary size edges (idxSize 1->2->3)
    d1c = desc([("s", "varchar(255)")])

    def dict_total(strings):
        return 1 + sum(len(s) + 1 for s in set(strings) if s)

    for target in (255, 256, 257, 65535, 65536, 65537):
        strings = []
        i = 0
        while dict_total(strings) + 12 < target:
            strings.append(b"k%09d" % i)   # 11 bytes incl. NUL
            i += 1
        rem = target - dict_total(strings)
        if rem >= 2:
            strings.append(b"z" * (rem - 1))
        assert dict_total(strings) == target, (target, dict_total(strings))
        add(f"d9_dict{target}", d1c, rows([[s] for s in strings]))

    # d10: numeric range edges
    d = desc([("u", "bigint(20) unsigned"), ("s", "bigint(20)")])
    for span in (254, 255, 256, 65535, 65536, 16777215, 16777216, 4294967295, 4294967296, 2**40, 2**48, 2**56 - 1, 2**56):
        add(f"d10_span{span}", d, rows([[b"1000", b"-5"], [b"%d" % (1000 + span), b"7"], [b"0", b"0"], [b"1000", b"-5"]]))
    add("d10_full", d, rows([[b"1", b"1"], [b"18446744073709551615", b"-1"], [b"1", b"-9223372036854775808"]]))

    # d11: wrong field count
    d = desc([("a", "varchar(8)"), ("b", "int(11)"), ("c", "text")])
    ok = [b"a", b"1", b"c"]
    add("d11_few_first", d, rows([[b"a", b"1"], ok, ok]), expect_rc=15)
    add("d11_many_mid", d, rows([ok, ok + [b"x"], ok]), expect_rc=15)
    add("d11_few_last", d, rows([ok, ok, [b"a"]]), expect_rc=15)
    add("d11_many_rows_bad_late", d, rows([ok] * 500 + [[b"a", b"1"]] + [ok] * 20), expect_rc=15)

    # d12: empty input / single row / single column
    add("d12_empty", d, b"")
    add("d12_only_blank", d, b"\n\n\n")
    add("d12_single_row", d, rows([ok]))
    add("d12_single_col", desc([("only", "varchar(8)")]), b"a\nb\na\n")
    add("d12_single_col_int", desc([("only", "int(11)")]), b"10\n20\n10\n")

    # d13: .desc.sql oddities
    d = (b"Field\tType\n" + desc([("f1", "float"), ("f2", "double"), ("f3", "timestamp"), ("f4", "mediumint(9)"),
                                   ("f5", " decimal(3,1)"), ("f6", "varchar(70000)"), ("f7", "char(0)"),
                                   ("f8", "enum('a','b')"), ("fieldx", "int(11)"), ("f9", "date")]))
    # NB: the column named "fieldx" is skipped by the strncasecmp("Field") rule -> 9 columns
    add("d13_desc_oddities", d, rows([[b"1.5", b"2.5", b"20190901", b"7", b"1.0", b"long", b"c0", b"a", b"2019-09-01"],
                                       [b"2", b"3", b"20190902", b"8", b"", b"long", b"", b"b", b"2019-09-02"]]))

    # trim mode (-t): trailing spaces stripped from every field before both passes
    d = desc([("a", "varchar(8)"), ("b", "int(11)"), ("c", "text"), ("d", "char(1)")])
    t = rows([[b"a  ", b"1 ", b"c", b"x"], [b"   ", b"2", b" c ", b" "], [b"a", b"1   ", b"c  \t".rstrip(b"\t"), b"y  "],
              [b"a" + BS + b" ", b"3", b"  ", b""]])
    add("trim_spaces", d, t, trim=True)
    add("trim_off_same_input", d, t)

    # mixed random table, moderate size, many repeats (exercises flags)
    cols = [("id", "int(11) unsigned"), ("url", "varchar(255)"), ("n", "smallint(5) unsigned"), ("t", "text"),
            ("ch", "char(1)"), ("dec", "decimal(24,12)"), ("neg", "int(11)"), ("e", "varchar(10)")]
    rng = random.Random(20190901)
    urls = [b"http://www.example.com/" + (b"%x" % rng.getrandbits(40)) for _ in range(50)]
    rs = []
    cur = [b"1", urls[0], b"5", b"t", b"a", b"1.5", b"-3", b""]
    for i in range(3000):
        if rng.random() < 0.9:
            cur[0] = b"%d" % (i + 1)
        if rng.random() < 0.3:
            cur[1] = rng.choice(urls)
        if rng.random() < 0.1:
            cur[2] = b"%d" % rng.randrange(0, 70000)
        if rng.random() < 0.05:
            cur[3] = rng.choice([b"", b"t", b"longer text value", b"x" * 100])
        if rng.random() < 0.2:
            cur[4] = rng.choice([b"", b"a", b"b", BS + b"t"])
        if rng.random() < 0.1:
            cur[5] = rng.choice([b"", b"1.5", b"0.000000000000", b"99.25"])
        if rng.random() < 0.1:
            cur[6] = b"%d" % rng.randrange(-1000, 1000)
        rs.append(list(cur))
    add("mixed_3000", desc(cols), rows(rs))

    # d14: strings that agree on long prefixes (URLs do): the dictionary order is decided 15 .. 130 bytes into the strings,
    # by a proper-prefix relation, by a high or low byte right behind the common part - the sort's prefix keys tie and its
    # in-memory compares (16 bytes a round, then words, then bytes) decide.  ~2 500 strings: several sort tiles.
    strings = []
    for k, L in enumerate((15, 16, 17, 19, 31, 32, 33, 47, 48, 49, 63, 64, 65, 100, 130)):
        stem = (b"http://www.example.com/%02d/" % k + b"path/" * 30)[:L]
        assert len(stem) == L
        for tail in (b"", b"a", b"b", b"\x01", b"\xff", b"aa", b"ab", b"a\x01", b"a" * 15, b"a" * 16, b"a" * 17, b"b" * 3 + b"\x7f",
                     b"\x80", b"\xfe\xff", b"zzzzzzz"):
            strings.append(stem + tail)
        for j in range(150):
            strings.append(stem + b"%c%05d" % (65 + j % 26, (j * 7919) % 100000))
    rng14 = random.Random(14)
    rng14.shuffle(strings)
    add("d14_long_common_prefixes", d1c, rows([[x] for x in strings]))
    return out


def _big_cases():
    """Cases too large for the loops that walk the whole corpus: run once each by dedicated tests."""
    out = []
    d1c = desc([("s", "varchar(255)")])
    # d9: dictionary total at 2^24 - 1 / 2^24 / 2^24 + 1 bytes: 3- vs 4-byte offsets (dictionary.cpp:62-73)
    for target in (16777215, 16777216, 16777217):
        n = (target - 1 - 16) // 11            # 11 bytes per "k%09d" + NUL; the last string takes up the slack
        total = 1 + 11 * n
        rem = target - total
        strings = [b"k%09d" % i for i in range(n)]
        if rem >= 2:
            strings.append(b"z" * (rem - 1))
        assert 1 + sum(len(x) + 1 for x in strings) == target
        out.append((f"d9_dict{target}", d1c, b"\n".join(strings) + b"\n", {}))
    return out


_BIG = None


def big_cases():
    global _BIG
    if _BIG is None:
        _BIG = _big_cases()
    return _BIG


def later_block_long_line():
    """d4: the long lines sit in LATER blocks (2 rows per block): longestLine is cumulative over the file,
    32768 (the interrupted row behind block 0 is read in full) -> 32768 -> 131072 -> 131072 (getnextrow.cpp:57-65,
    ConvertToZDW.cpp:965)."""
    two_text = desc([("a", "text"), ("b", "text")])
    r = [b"s\tt", b"u\tv", b"x" * 20000 + b"\ty", b"p\tq", b"m\tn", b"z" * 70000 + b"\tw", b"e\tf", b"g\th"]
    return two_text, b"\n".join(r) + b"\n", 2


_CACHE = None


def cases():
    global _CACHE
    if _CACHE is None:
        _CACHE = _cases()
    return _CACHE


def case_ids():
    return [c[0] for c in cases()]
