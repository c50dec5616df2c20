"""The drop-in boundary itself (no GPU needed): libzdw_b200.so loads, exports every function include/zdw_b200.h
declares, reports the header's ABI version, refuses to create a context without a device (there is no CPU fallback),
and the ctypes mirror in zdw_b200/capi.py lays its structs out exactly like a C compiler lays out the header's."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "zdw_b200.h"
sys.path.insert(0, str(ROOT))


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(zdwb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_function():
    from zdw_b200 import capi
    lib = C.CDLL(str(capi.lib_path()))
    names = declared_functions()
    assert {"zdwb_ctx_create", "zdwb_ctx_destroy", "zdwb_encode_block", "zdwb_decode_block", "zdwb_last_error"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_matches_header():
    from zdw_b200 import capi
    lib = capi.load_library()
    want = int(re.search(r"#define\s+ZDWB_ABI_VERSION\s+(\d+)", HEADER.read_text()).group(1))
    assert lib.zdwb_abi_version() == want


def test_no_device_no_context():
    """Without a GPU the library says so; nothing in the product computes on the CPU instead."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zdw_b200 import capi
    lib = capi.load_library()
    h = C.c_void_p()
    rc = lib.zdwb_ctx_create(0, 0, C.byref(h))
    assert rc != 0 and not h.value
    with pytest.raises(capi.ZdwError):
        capi.Context(0)


def test_ctypes_mirror_matches_the_header_layout(tmp_path):
    from zdw_b200 import capi
    pairs = [("zdwb_schema", capi._Schema), ("zdwb_encode_opts", capi._EncOpts), ("zdwb_block_out", capi._BlockOut),
             ("zdwb_fill", capi._Fill), ("zdwb_decode_opts", capi._DecOpts), ("zdwb_rows_out", capi._RowsOut)]
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{HEADER}"', "int main(void) {"]
    fields = {}
    for cname, _ in pairs:
        body = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + cname + r"\s*;", text).group(1)
        names = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?\s*;", body)
        fields[cname] = names
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for f in names:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    for (cname, mirror), line in zip(pairs, out):
        nums = [int(x) for x in line.split()[1:]]
        assert C.sizeof(mirror) == nums[0], cname
        offs = [getattr(mirror, f[0]).offset for f in mirror._fields_]
        assert offs == nums[1:], (cname, fields[cname], [f[0] for f in mirror._fields_])
