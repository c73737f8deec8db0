"""The C-ABI library builds, loads and exports every symbol include/tvbf.h declares (CPU only:
no compute call is made)."""

import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tvbingefriend_recommendation_service_b200 import _lib, build

    build.build()
    return _lib.load()


def declared_functions():
    text = (ROOT / "include" / "tvbf.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tvbf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("tvbf_hybrid_topk", "tvbf_exact_rows", "tvbf_prep_csr_to_operand", "tvbf_cosine_matrix_f64",
                 "tvbf_matrix_stats_f64", "tvbf_matrix_rows_topk", "tvbf_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from tvbingefriend_recommendation_service_b200 import _lib

    names = declared_functions()
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    for name in names:
        assert getattr(lib, name) is not None


def test_version_and_error_string(lib):
    assert lib.tvbf_version() == 100
    assert isinstance(lib.tvbf_last_error(), bytes)


def test_struct_layout_matches_header(lib):
    """ctypes mirrors must have the C layout (sizes computed by the compiler for the header)."""
    import ctypes as C
    import subprocess
    import tempfile

    from tvbingefriend_recommendation_service_b200 import _lib

    src = '#include "tvbf.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(tvbf_features), sizeof(tvbf_params), sizeof(tvbf_topk_out));return 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        c = Path(tmp) / "s.c"
        c.write_text(src)
        exe = Path(tmp) / "s"
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
        sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.Features), C.sizeof(_lib.Params), C.sizeof(_lib.TopKOut)]


def test_sass_is_blackwell_native():
    """The hot kernel must contain tcgen05 MMA, TMEM loads and TMA loads (SASS mnemonics)."""
    import shutil
    import subprocess

    from tvbingefriend_recommendation_service_b200 import build

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(build.build())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_no_gpu_means_loud_failure(lib):
    import torch

    from tvbingefriend_recommendation_service_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.TvbfError):
        _lib.require_device()
    from tvbingefriend_recommendation_service_b200.ml.similarity_computer import SimilarityComputer
    import numpy as np

    with pytest.raises(_lib.TvbfError):
        SimilarityComputer().compute_genre_similarity(np.eye(3))
