"""Build ``libtvbf.so`` (the sm_100a CUDA library behind the C ABI in ``include/tvbf.h``) in-tree.

    python -m tvbingefriend_recommendation_service_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting ``.so`` is git-ignored but travels
to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libtvbf.so"
STAMP = PKG / "csrc" / ".build_stamp"
SOURCES = ["api.cu", "prep.cu", "hybrid_topk.cu", "rescore.cu", "matrix.cu", "moments.cu"]
HEADERS = [CSRC / "common.cuh", CSRC / "internal.cuh", PKG.parent / "include" / "tvbf.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("TVBF_EXTRA_NVCC_FLAGS", "").split()   # experiments only (e.g. -DTVBF_MBAR_HINT_NS=1000)
# The exact fp64 scorers must round like numpy (one rounding per product and per sum): no FMA
# contraction anywhere in rescore.cu.  (The hybrid expression itself is written with explicit
# __dmul_rn / __dadd_rn in common.cuh: hybrid_rn.)
PER_FILE_FLAGS = {"rescore.cu": ["-fmad=false"]}


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: libtvbf.so cannot be built")
    return cand


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(sorted(PER_FILE_FLAGS.items())).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every translation unit and link the shared library; returns its path."""
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *PER_FILE_FLAGS.get(src, []), "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                 text=True)))
    objs = []
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src}\n{out}")
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(str(obj))
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(path, os.path.getsize(path), "bytes")
