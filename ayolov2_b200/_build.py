"""In-tree build of libay2.so (CUDA kernels + C-ABI) for sm_100a with nvcc.

The shared object is written next to the sources (ayolov2_b200/libay2.so) so that it travels with the
repo snapshot to the GPU box; it is git-ignored. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libay2.so"
HASH_MARK = b"AY2_SOURCE_HASH="

SOURCES = ["capi.cu", "conv_tc.cu", "conv_chain.cu", "conv_wgrad.cu", "pointwise.cu", "precise.cu", "train_pointwise.cu", "nms.cu", "nms_variants.cu", "loss.cu", "val_match.cu", "letterbox.cu", "pseudo_labels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libay2.so cannot be built")


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")))
    files.append(PKG_DIR.parent / "include" / "ay2.h")
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def embedded_hash(path: Path = LIB_PATH) -> str:
    """The source hash compiled into libay2.so (csrc/capi.cu, `ay2_source_hash()`), read from the file's bytes so that
    the check needs neither a loadable CUDA runtime nor a side file that could travel separately from the binary."""
    if not path.exists():
        return ""
    data = path.read_bytes()
    i = data.find(HASH_MARK)
    if i < 0:
        return ""
    return data[i + len(HASH_MARK):i + len(HASH_MARK) + 64].decode("ascii", "replace")


def is_current() -> bool:
    return embedded_hash() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every translation unit to an object and link libay2.so. Rebuilds only when the hash embedded in the
    existing binary differs from the hash of the sources (a stale .so after a checkout is rebuilt, never silently bound)."""
    want = _source_hash()
    if not force and embedded_hash() == want:
        return LIB_PATH
    nvcc = _nvcc()
    objdir = PKG_DIR / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        sp = CSRC / src
        if not sp.exists():
            continue
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(sp), "-o", str(obj)]
        if src == "capi.cu":
            cmd.insert(1, f"-DAY2_SOURCE_HASH_STR=\"{want}\"")
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    (objdir / "build.log").write_text("\n".join(log))
    assert embedded_hash() == want, "libay2.so does not carry the source hash it was built from"
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
