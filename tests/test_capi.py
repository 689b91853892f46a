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
