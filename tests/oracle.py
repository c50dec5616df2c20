"""ctypes binding of the plain-C oracle (oracle/libzdw_oracle.so) + helpers to drive oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (zdw_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import lzma
import os
import shutil
import subprocess
import tempfile
from dataclasses import dataclass
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
GOLDEN = ROOT / "tests" / "golden"

TEXT_LIKE = {0, 1, 2, 3, 11, 16, 17, 18}


def build_oracle() -> Path:
    so = ORACLE_DIR / "libzdw_oracle.so"
    src = ORACLE_DIR / "zdw_oracle.c"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, (ORACLE_DIR / "zdw_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "port"], check=True, capture_output=True)
    return so


class _Schema(C.Structure):
    _fields_ = [("ncols", C.c_uint32), ("names", C.POINTER(C.c_char_p)), ("types", C.POINTER(C.c_uint8)),
                ("charsize", C.POINTER(C.c_uint16))]


class _EncOpts(C.Structure):
    _fields_ = [("trim", C.c_int), ("rows_per_block", C.c_uint32), ("nmeta", C.c_uint32),
                ("keys", C.POINTER(C.c_char_p)), ("vals", C.POINTER(C.c_char_p)),
                ("nplan", C.c_uint32), ("plan_rows", C.POINTER(C.c_uint32)), ("plan_spill", C.POINTER(C.c_uint32)),
                ("heap_blocks", C.c_uint32)]


class _EncInfo(C.Structure):
    _fields_ = [("total_rows", C.c_uint64), ("nblocks", C.c_uint32), ("longest_line", C.c_uint32),
                ("bad_row", C.c_uint32), ("dict_entries", C.c_uint64)]


class _DecOpts(C.Structure):
    _fields_ = [("out_col", C.POINTER(C.c_int)), ("n_out", C.c_uint32), ("sep", C.c_char)]


class _DecInfo(C.Structure):
    _fields_ = [("version", C.c_uint16), ("ncols", C.c_uint32), ("total_rows", C.c_uint64), ("nblocks", C.c_uint32),
                ("line_length", C.c_uint32), ("consumed", C.c_size_t), ("block_rows", C.c_uint32 * 16),
                ("block_offset", C.c_uint64 * 16)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build_oracle()))
        L.zo_parse_desc.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Schema)]
        L.zo_schema_free.argtypes = [C.POINTER(_Schema)]
        L.zo_encode_file.argtypes = [C.POINTER(_Schema), C.c_char_p, C.c_size_t, C.POINTER(_EncOpts),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(_EncInfo)]
        L.zo_decode_file.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_DecOpts), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t), C.POINTER(_DecInfo)]
        L.zo_read_file_header.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Schema), C.POINTER(C.c_uint16),
                                          C.POINTER(C.c_size_t)]
        L.zo_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


@dataclass
class Schema:
    names: list
    types: list
    charsize: list

    @property
    def ncols(self):
        return len(self.types)


def parse_desc(text: bytes) -> Schema:
    s = _Schema()
    rc = lib().zo_parse_desc(text, len(text), C.byref(s))
    if rc:
        raise ValueError(f"desc parse error {rc}")
    out = Schema([s.names[i].decode("latin1") for i in range(s.ncols)], [s.types[i] for i in range(s.ncols)],
                 [s.charsize[i] for i in range(s.ncols)])
    lib().zo_schema_free(C.byref(s))
    return out


def _c_schema(sch: Schema):
    n = sch.ncols
    names = (C.c_char_p * max(n, 1))(*[x.encode("latin1") for x in sch.names])
    types = (C.c_uint8 * max(n, 1))(*sch.types)
    cs = (C.c_uint16 * max(n, 1))(*sch.charsize)
    s = _Schema(n, C.cast(names, C.POINTER(C.c_char_p)), C.cast(types, C.POINTER(C.c_uint8)),
                C.cast(cs, C.POINTER(C.c_uint16)))
    s._keep = (names, types, cs)
    return s


@dataclass
class EncodeResult:
    rc: int
    data: bytes
    total_rows: int
    nblocks: int
    longest_line: int
    bad_row: int
    dict_entries: int


def encode(sch: Schema, tsv: bytes, trim: bool = False, rows_per_block: int = 0, metadata: dict | None = None,
           plan=None, heap_blocks: int = 0) -> EncodeResult:
    """plan: [(rows, spill_columns), ...] - explicit block boundaries with the reference's interrupted-row spill.
    heap_blocks: K - the reference's own cut: a block ends on the insert that opens its K-th 64 MiB string-heap block."""
    s = _c_schema(sch)
    md = sorted((metadata or {}).items())
    keys = (C.c_char_p * max(len(md), 1))(*[k.encode() for k, _ in md])
    vals = (C.c_char_p * max(len(md), 1))(*[v.encode() for _, v in md])
    plan = list(plan or [])
    prow = (C.c_uint32 * max(len(plan), 1))(*[int(r) for r, _ in plan])
    pspill = (C.c_uint32 * max(len(plan), 1))(*[int(c) for _, c in plan])
    o = _EncOpts(int(trim), rows_per_block, len(md), C.cast(keys, C.POINTER(C.c_char_p)), C.cast(vals, C.POINTER(C.c_char_p)),
                 len(plan), C.cast(prow, C.POINTER(C.c_uint32)), C.cast(pspill, C.POINTER(C.c_uint32)), int(heap_blocks))
    out = C.c_void_p()
    n = C.c_size_t()
    info = _EncInfo()
    rc = lib().zo_encode_file(C.byref(s), tsv, len(tsv), C.byref(o), C.byref(out), C.byref(n), C.byref(info))
    data = C.string_at(out, n.value) if out.value else b""
    if out.value:
        lib().zo_free(out)
    return EncodeResult(rc, data, info.total_rows, info.nblocks, info.longest_line, info.bad_row, info.dict_entries)


@dataclass
class DecodeResult:
    rc: int
    tsv: bytes
    version: int
    ncols: int
    total_rows: int
    nblocks: int
    line_length: int
    consumed: int
    block_rows: list = None     # numRows of the first 16 blocks
    block_offset: list = None   # offsets of their block headers in the image


def decode(zdw: bytes, out_col: list | None = None, n_out: int = 0, sep: bytes = b"\t") -> DecodeResult:
    arr = None
    if out_col is not None:
        arr = (C.c_int * len(out_col))(*out_col)
    o = _DecOpts(C.cast(arr, C.POINTER(C.c_int)) if arr is not None else None, n_out, sep)
    out = C.c_void_p()
    n = C.c_size_t()
    info = _DecInfo()
    rc = lib().zo_decode_file(zdw, len(zdw), C.byref(o), C.byref(out), C.byref(n), C.byref(info))
    data = C.string_at(out, n.value) if out.value else b""
    if out.value:
        lib().zo_free(out)
    nb = min(info.nblocks, 16)
    return DecodeResult(rc, data, info.version, info.ncols, info.total_rows, info.nblocks, info.line_length, info.consumed,
                        list(info.block_rows[:nb]), list(info.block_offset[:nb]))


def read_header(zdw: bytes):
    s = _Schema()
    ver = C.c_uint16()
    hl = C.c_size_t()
    rc = lib().zo_read_file_header(zdw, len(zdw), C.byref(s), C.byref(ver), C.byref(hl))
    if rc:
        raise ValueError(f"header error {rc}")
    sch = Schema([s.names[i].decode("latin1") for i in range(s.ncols)], [s.types[i] for i in range(s.ncols)],
                 [s.charsize[i] for i in range(s.ncols)])
    lib().zo_schema_free(C.byref(s))
    return sch, ver.value, hl.value


# ----------------------------------------------------------------------------- golden fixtures

def golden(name: str) -> bytes:
    p = GOLDEN / name
    if p.exists():
        return p.read_bytes()
    px = GOLDEN / (name + ".xz")
    return lzma.decompress(px.read_bytes())


def golden_to_v11(zdw: bytes) -> bytes:
    """v9/v10 golden -> the v11 image the current encoder writes:
    version word replaced and a zero 4-byte metadata length inserted."""
    ver = int.from_bytes(zdw[:2], "little")
    assert ver in (9, 10)
    return (11).to_bytes(2, "little") + (0).to_bytes(4, "little") + zdw[2:]


# ----------------------------------------------------------------------------- compiled reference

def have_ref() -> bool:
    return (REF_DIR / "convertDWfile").exists() and (REF_DIR / "unconvertDWfile").exists()


def _ref_env():
    env = dict(os.environ)
    env["PATH"] = f"{REF_DIR / 'nocomp'}:{env.get('PATH', '')}"
    return env


def ref_encode(tsv: bytes, desc: bytes, args: list | None = None, timeout: int = 120, metadata_file: bytes | None = None):
    """Run the unmodified reference convertDWfile (compressor stage replaced by a pass-through).
    Returns (exit_code, zdw_bytes_or_None, stdout+stderr)."""
    d = tempfile.mkdtemp(prefix="zdwref_")
    try:
        (Path(d) / "x.sql").write_bytes(tsv)
        (Path(d) / "x.desc.sql").write_bytes(desc)
        if metadata_file is not None:
            (Path(d) / "x.metadata").write_bytes(metadata_file)
        p = subprocess.run([str(REF_DIR / "convertDWfile"), *(args or []), "x.sql"], cwd=d, env=_ref_env(),
                           capture_output=True, timeout=timeout)
        f = Path(d) / "x.zdw.gz"
        return p.returncode, (f.read_bytes() if f.exists() else None), (p.stdout + p.stderr).decode("latin1")
    finally:
        shutil.rmtree(d, ignore_errors=True)


def ref_decode(zdw: bytes, args: list | None = None, timeout: int = 120):
    """Run the unmodified reference unconvertDWfile to stdout. Returns (exit_code, tsv, stderr)."""
    d = tempfile.mkdtemp(prefix="zdwref_")
    try:
        (Path(d) / "x.zdw").write_bytes(zdw)
        p = subprocess.run([str(REF_DIR / "unconvertDWfile"), "-q", *(args or []), "-", "x.zdw"], cwd=d, env=_ref_env(),
                           capture_output=True, timeout=timeout)
        return p.returncode, p.stdout, p.stderr.decode("latin1")
    finally:
        shutil.rmtree(d, ignore_errors=True)


def ref_api_decode(zdw: bytes, args: list | None = None, timeout: int = 120):
    d = tempfile.mkdtemp(prefix="zdwref_")
    try:
        (Path(d) / "x.zdw").write_bytes(zdw)
        p = subprocess.run([str(REF_DIR / "test_unconvert_api"), *(args or []), "x.zdw"], cwd=d, env=_ref_env(),
                           capture_output=True, timeout=timeout)
        return p.returncode, p.stdout, p.stderr.decode("latin1")
    finally:
        shutil.rmtree(d, ignore_errors=True)
