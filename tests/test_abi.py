"""The C-ABI library: builds for sm_100a, loads, exports exactly what include/*.h declares.
No compute calls here (no GPU in the CPU suite)."""
import ctypes
import re
from pathlib import Path

import pytest

from haptools_b200 import _cuda
from haptools_b200._cuda import build as cuda_build

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "haptools_cuda.h"


@pytest.fixture(scope="module")
def lib():
    cuda_build.build()  # no-op when up to date
    return _cuda.lib()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ht_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(_cuda.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(str(_cuda.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} is declared in {HEADER.name} but not exported"


def test_abi_version_matches_header(lib):
    m = re.search(r"#define HT_ABI_VERSION (\d+)", HEADER.read_text())
    assert lib.ht_version() == int(m.group(1)) == _cuda.ABI_VERSION


def test_constants_match_header():
    text = HEADER.read_text()
    for name in ("HT_PACK_AUTO", "HT_PACK_GENERIC", "HT_PACK_VARIANT_MAJOR", "HT_PACK_SAMPLE_MAJOR",
                 "HT_PACK_SAMPLE_MAJOR_BIALLELIC",
                 "HT_COPY_H2D", "HT_COPY_D2H", "HT_COPY_D2D", "HT_ERR_BAD_ARG", "HT_ERR_CUDA", "HT_ERR_UNSUPPORTED"):
        m = re.search(rf"#define {name} (\d+)", text)
        assert m and int(m.group(1)) == getattr(_cuda, name), name
    assert int(re.search(r"#define HT_TILE_SAMPLES (\d+)", text).group(1)) == _cuda.TILE_SAMPLES


def test_size_helpers(lib):
    assert lib.ht_num_tiles(0) == 0 and lib.ht_num_tiles(1) == 1
    assert lib.ht_num_tiles(1024) == 1 and lib.ht_num_tiles(1025) == 2
    assert lib.ht_plane_bytes(2504, 10) == 3 * 10 * 256
    assert lib.ht_plane_bytes(0, 10) == 0 and lib.ht_plane_bytes(10, 0) == 0


def test_argument_errors_do_not_need_a_gpu(lib):
    # argument validation happens before any CUDA call
    rc = lib.ht_transform_u8(None, -1, 0, None, None, None, None, 0, None, 0, None, None)
    assert rc == _cuda.HT_ERR_BAD_ARG and b"negative" in lib.ht_last_error()
    with pytest.raises(ValueError):
        _cuda.check(rc, "ht_transform_u8")
    rc = lib.ht_pack_u8(None, 0, 0, 0, None, 0, 0, 0, 5, None, None, None, 0, None)
    assert rc == _cuda.HT_ERR_BAD_ARG and b"tables" in lib.ht_last_error()
    tables = _cuda.PackTables(8, 3, 3, None, None, None, None, None, None, None, 0)
    rc = lib.ht_pack_u8(None, 0, 0, 0, None, 0, 0, 0, 5, ctypes.byref(tables), None, None, 0, None)
    assert rc == _cuda.HT_ERR_BAD_ARG
    rc = lib.ht_copy2d(None, 0, None, 0, 4, 4, 1, None)
    assert rc == _cuda.HT_ERR_BAD_ARG


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never fall back."""
    import numpy as np
    import torch

    from haptools_b200 import engine
    from haptools_b200.plan import plan_from_refs

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    plan = plan_from_refs(np.array([0]), np.array([1]), np.array([0, 1]), 3)
    with pytest.raises(_cuda.HaptoolsCudaError):
        engine.transform_host(np.zeros((4, 3, 2), dtype=np.uint8), plan)


def test_product_never_imports_the_oracle():
    for path in (REPO / "haptools_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
