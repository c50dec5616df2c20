"""ctypes binding of the libzdw_b200 C ABI (include/zdw_b200.h).

Fails loudly (ImportError / ZdwError) when the CUDA library is missing or no GPU is present:
there is deliberately no CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path

_HERE = Path(__file__).resolve().parent

STATUS_NAMES = {0: "OK", 1: "ERR_CUDA", 2: "ERR_OOM", 3: "ERR_WRONG_COLUMNS", 4: "ERR_BAD_ARG", 5: "ERR_UNSUPPORTED",
                6: "ERR_CORRUPT", 7: "ERR_TRUNCATED", 8: "ERR_ROW_COUNT", 9: "ERR_NO_DEVICE"}

EXPORTED_SYMBOLS = [
    "zdwb_abi_version", "zdwb_device_count", "zdwb_ctx_create", "zdwb_ctx_destroy", "zdwb_last_error", "zdwb_ctx_set_stream",
    "zdwb_ctx_set_tuning", "zdwb_ctx_kernel_launches", "zdwb_ctx_kernel_times", "zdwb_encode_block", "zdwb_decode_block", "zdwb_host_alloc",
    "zdwb_host_free", "zdwb_fd_to_device", "zdwb_device_to_fd",
]


class ZdwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


def lib_path() -> Path:
    # ZDWB_LIB: developer override used by tools/ to A/B two builds of the same library (never a fallback)
    override = os.environ.get("ZDWB_LIB")
    return Path(override) if override else _HERE / "libzdw_b200.so"


class _Schema(C.Structure):
    _fields_ = [("ncols", C.c_uint32), ("types", C.POINTER(C.c_uint8))]


class _EncOpts(C.Structure):
    _fields_ = [("trim", C.c_int32), ("input_on_device", C.c_int32), ("output_on_device", C.c_int32),
                ("more_input_follows", C.c_int32), ("prev_longest_line", C.c_uint32), ("spill_cols", C.c_uint32),
                ("max_rows", C.c_uint64), ("heap_blocks", C.c_uint32), ("reserved_e", C.c_uint32)]


class _BlockOut(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("len", C.c_size_t), ("nrows", C.c_uint32), ("longest_line", C.c_uint32),
                ("tsv_consumed", C.c_uint64), ("rows_in_buffer", C.c_uint64), ("bad_row", C.c_uint32),
                ("ncols_used", C.c_uint32), ("dict_entries", C.c_uint64), ("dict_bytes", C.c_uint64),
                ("dict_index_size", C.c_uint32), ("reserved", C.c_uint32)]


class _Fill(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("len", C.c_uint32), ("text", C.c_char_p)]


class _DecOpts(C.Structure):
    _fields_ = [("input_on_device", C.c_int32), ("output_on_device", C.c_int32), ("want_row_offsets", C.c_int32),
                ("at_end_of_file", C.c_int32), ("separator", C.c_uint8), ("reserved", C.c_uint8 * 7),
                ("out_col", C.POINTER(C.c_int32)), ("n_out", C.c_uint32), ("n_fills", C.c_uint32),
                ("fills", C.POINTER(_Fill)), ("rownum_pos", C.c_int32), ("validate_only", C.c_int32),
                ("first_row_number", C.c_uint64), ("want_flag_counts", C.c_int32), ("skim_only", C.c_int32)]


class _RowsOut(C.Structure):
    _fields_ = [("tsv", C.c_void_p), ("len", C.c_size_t), ("row_off", C.c_void_p), ("nrows", C.c_uint32),
                ("line_length", C.c_uint32), ("is_last", C.c_uint8), ("reserved", C.c_uint8 * 7),
                ("consumed", C.c_uint64), ("dict_bytes", C.c_uint64), ("ncols_used", C.c_uint32),
                ("reserved2", C.c_uint32), ("flag_counts", C.c_void_p)]


_lib = None


def _copy_out(ptr: int, n: int) -> bytes:
    """n bytes at a raw host pointer (ctypes.string_at takes a C int: it fails beyond 2 GiB)."""
    if not n:
        return b""
    return bytes((C.c_uint8 * n).from_address(ptr))


def load_library():
    """Loads libzdw_b200.so; raises ImportError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise ImportError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(zdw_b200 has no CPU fallback)")
    L = C.CDLL(str(p))
    L.zdwb_abi_version.restype = C.c_int
    L.zdwb_device_count.restype = C.c_int
    L.zdwb_ctx_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.zdwb_ctx_destroy.argtypes = [C.c_void_p]
    L.zdwb_ctx_destroy.restype = None
    L.zdwb_last_error.argtypes = [C.c_void_p]
    L.zdwb_last_error.restype = C.c_char_p
    L.zdwb_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.zdwb_ctx_set_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
    L.zdwb_ctx_kernel_launches.argtypes = [C.c_void_p]
    L.zdwb_ctx_kernel_launches.restype = C.c_ulonglong
    L.zdwb_ctx_kernel_times.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.zdwb_ctx_kernel_times.restype = C.c_size_t
    L.zdwb_encode_block.argtypes = [C.c_void_p, C.POINTER(_Schema), C.c_void_p, C.c_size_t, C.POINTER(_EncOpts),
                                    C.POINTER(_BlockOut)]
    L.zdwb_decode_block.argtypes = [C.c_void_p, C.POINTER(_Schema), C.c_void_p, C.c_size_t, C.POINTER(_DecOpts),
                                    C.POINTER(_RowsOut)]
    L.zdwb_host_alloc.argtypes = [C.c_size_t]
    L.zdwb_host_alloc.restype = C.c_void_p
    L.zdwb_host_free.argtypes = [C.c_void_p]
    L.zdwb_host_free.restype = None
    L.zdwb_fd_to_device.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.POINTER(C.c_void_p)]
    L.zdwb_device_to_fd.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_longlong]
    _lib = L
    return L


@dataclass
class EncodedBlock:
    data: bytes | None      # block bytes (None when left on the device)
    dev_ptr: int            # device pointer when output_on_device
    length: int
    nrows: int
    longest_line: int
    tsv_consumed: int
    rows_in_buffer: int
    ncols_used: int
    dict_entries: int
    dict_bytes: int
    dict_index_size: int


@dataclass
class DecodedBlock:
    tsv: bytes | None
    dev_ptr: int
    length: int
    row_off: list | None
    nrows: int
    line_length: int
    is_last: bool
    consumed: int
    dict_bytes: int
    ncols_used: int
    flag_counts: list | None = None


class Context:
    """One GPU context (zdwb_ctx).  Not thread-safe; one per host thread / GPU."""

    def __init__(self, device: int = 0, workspace_hint: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.zdwb_ctx_create(device, workspace_hint, C.byref(h))
        if rc != 0:
            raise ZdwError(rc, "zdwb_ctx_create failed (is a CUDA device visible? zdw_b200 has no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.zdwb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def last_error(self) -> str:
        return self._L.zdwb_last_error(self._h).decode("latin1")

    def set_stream(self, cuda_stream: int | None):
        rc = self._L.zdwb_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0))
        if rc:
            raise ZdwError(rc, self.last_error())

    def set_tuning(self, name: str, value: int):
        rc = self._L.zdwb_ctx_set_tuning(self._h, name.encode(), value)
        if rc:
            raise ZdwError(rc, f"unknown tuning knob {name}")

    def kernel_launches(self) -> int:
        return int(self._L.zdwb_ctx_kernel_launches(self._h))

    def kernel_times(self) -> dict:
        """{kernel name: (launches, total_ms)} since the last call (needs set_tuning("kernel_timing", 1))."""
        buf = C.create_string_buffer(1 << 16)
        self._L.zdwb_ctx_kernel_times(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    # ------------------------------------------------------------------ file descriptor <-> device
    def fd_to_device(self, fd: int, offset: int, n: int) -> int:
        """`n` bytes of fd from `offset` (< 0: read() at its position) to a device buffer of the context; returns its address."""
        dev = C.c_void_p()
        rc = self._L.zdwb_fd_to_device(self._h, fd, offset, n, C.byref(dev))
        if rc:
            raise ZdwError(rc, self.last_error())
        return int(dev.value or 0)

    def device_to_fd(self, dev_ptr: int, n: int, fd: int, offset: int):
        """`n` device bytes to fd at `offset` (< 0: write() at its position)."""
        rc = self._L.zdwb_device_to_fd(self._h, C.c_void_p(dev_ptr), n, fd, offset)
        if rc:
            raise ZdwError(rc, self.last_error())

    # ------------------------------------------------------------------ encode
    def encode_block(self, types, tsv, n: int | None = None, *, trim=False, input_on_device=False,
                     output_on_device=False, prev_longest_line=0, max_rows=0, more_input_follows=False,
                     spill_cols=0, heap_blocks=0) -> EncodedBlock:
        """tsv: bytes-like (host) or an int device pointer (input_on_device=True, n required)."""
        tarr = (C.c_uint8 * max(len(types), 1))(*types)
        sch = _Schema(len(types), C.cast(tarr, C.POINTER(C.c_uint8)))
        if input_on_device:
            ptr = C.c_void_p(int(tsv))
            assert n is not None
        else:
            keep = tsv if isinstance(tsv, (bytes, bytearray)) else bytes(tsv)
            n = len(keep) if n is None else n
            ptr = C.cast(C.c_char_p(bytes(keep)), C.c_void_p) if not isinstance(keep, bytes) else C.cast(C.c_char_p(keep), C.c_void_p)
        o = _EncOpts(int(trim), int(input_on_device), int(output_on_device), int(more_input_follows), prev_longest_line,
                     int(spill_cols), max_rows, int(heap_blocks), 0)
        out = _BlockOut()
        rc = self._L.zdwb_encode_block(self._h, C.byref(sch), ptr, n, C.byref(o), C.byref(out))
        if rc:
            e = ZdwError(rc, self.last_error())
            e.bad_row = out.bad_row
            raise e
        data = None
        if not output_on_device:
            data = _copy_out(out.bytes, out.len)
        return EncodedBlock(data, int(out.bytes or 0), out.len, out.nrows, out.longest_line, out.tsv_consumed,
                            out.rows_in_buffer, out.ncols_used, out.dict_entries, out.dict_bytes, out.dict_index_size)

    # ------------------------------------------------------------------ decode
    def decode_block(self, types, zdw, avail: int | None = None, *, input_on_device=False, output_on_device=False,
                     want_row_offsets=False, at_end_of_file=True, separator=b"\t", out_col=None, n_out=0, fills=None,
                     rownum_pos=-1, first_row_number=1, validate_only=False, want_flag_counts=False,
                     skim_only=False) -> DecodedBlock:
        """fills: {output position: bytes} constant texts for positions no file column maps to."""
        tarr = (C.c_uint8 * max(len(types), 1))(*types)
        sch = _Schema(len(types), C.cast(tarr, C.POINTER(C.c_uint8)))
        if input_on_device:
            ptr = C.c_void_p(int(zdw))
            assert avail is not None
        else:
            keep = bytes(zdw)
            avail = len(keep) if avail is None else avail
            ptr = C.cast(C.c_char_p(keep), C.c_void_p)
        oc = None
        if out_col is not None:
            oc = (C.c_int32 * len(out_col))(*out_col)
        fl = None
        if fills:
            fl = (_Fill * len(fills))(*[_Fill(int(k), len(v), v) for k, v in fills.items()])
        o = _DecOpts(int(input_on_device), int(output_on_device), int(want_row_offsets), int(at_end_of_file),
                     separator[0], (C.c_uint8 * 7)(), C.cast(oc, C.POINTER(C.c_int32)) if oc is not None else None,
                     n_out, len(fills) if fills else 0, C.cast(fl, C.POINTER(_Fill)) if fl is not None else None,
                     rownum_pos, int(validate_only), first_row_number, int(want_flag_counts), int(skim_only))
        out = _RowsOut()
        rc = self._L.zdwb_decode_block(self._h, C.byref(sch), ptr, avail, C.byref(o), C.byref(out))
        if rc:
            raise ZdwError(rc, self.last_error())
        tsv = None
        offs = None
        if not output_on_device:
            tsv = _copy_out(out.tsv, out.len)
            if want_row_offsets and out.row_off:
                arr = (C.c_uint64 * (out.nrows + 1)).from_address(out.row_off)
                offs = list(arr)
        counts = None
        if want_flag_counts and out.flag_counts:
            counts = list((C.c_uint64 * out.ncols_used).from_address(out.flag_counts))
        return DecodedBlock(tsv, int(out.tsv or 0), out.len, offs, out.nrows, out.line_length, bool(out.is_last),
                            out.consumed, out.dict_bytes, out.ncols_used, counts)
