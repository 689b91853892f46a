"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/ay2.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ay2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ay2_[a-z0-9_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as g
    from ayolov2_b200 import _lib

    g.build()
    lib = ctypes.CDLL(str(_lib.lib_path()))
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ay2.h but not exported"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes prototypes out of sync with the header"
    assert _lib.load().ay2_version() >= 100


def test_struct_layouts_match_header():
    from ayolov2_b200 import _lib

    assert ctypes.sizeof(_lib.ConvDesc) == 112 and _lib.ConvDesc.in2.offset == 104  # 25 int32, 4 bytes padding, pointer
    assert ctypes.sizeof(_lib.ChainDesc) == 17 * 4
    assert ctypes.sizeof(_lib.NmsParams) == 48
    assert ctypes.sizeof(_lib.LossParams) == 120 and _lib.LossParams.fl_gamma.offset == 112  # 5 + 10 + 5 + 8 + 2 four-byte fields
    assert _lib.NmsParams.iou_thres.offset == 0 and _lib.NmsParams.batch.offset == 16


def test_sass_uses_blackwell_tensor_path():
    """The conv kernel must be tcgen05 / TMA code (UTCHMMA, UTMALDG, UTMASTG, LDTM in SASS), not mma.sync."""
    import shutil
    import subprocess

    from ayolov2_b200 import _lib

    if not shutil.which("cuobjdump"):
        return
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.lib_path())], stdout=subprocess.PIPE, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnem in sass, mnem
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_input_side_records_match_the_c_structs(tmp_path):
    """The host writes ay2_letterbox_image / ay2_load_resize_image records as numpy structured arrays: their field offsets and
    sizes must be what a C compiler makes of include/ay2.h (compiled here with gcc; plain C, no CUDA needed)."""
    import shutil
    import subprocess

    import pytest

    from ayolov2_b200 import data_loader as dl

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"ay2_letterbox_image": dl._REC, "ay2_load_resize_image": dl._LOAD_REC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ay2.h"', "int main(void) {"]
    for name, dt in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field in dt.names:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout
    seen = 0
    for line in out.splitlines():
        name, field, value = line.split()
        dt = structs[name]
        if field == "size":
            assert dt.itemsize == int(value), (name, dt.itemsize, value)
        else:
            assert dt.fields[field][1] == int(value), (name, field, dt.fields[field][1], value)
        seen += 1
    assert seen == sum(len(dt.names) + 1 for dt in structs.values())
