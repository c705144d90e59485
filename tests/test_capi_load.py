"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/fle_b200.h declares.  No compute call is made (there is no GPU here, and the library
has no CPU path: context creation must fail with FLE_E_NODEVICE)."""
import os
import re

import pytest

from conftest import ROOT, _has_gpu


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fle_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib_path):
    from fle_b200 import capi
    L = capi.load_library()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in fle_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names, set(capi.EXPORTS) ^ set(names)
    assert L.fle_version() == 100
    assert L.fle_encoded_bytes(100, 1) == 16 and L.fle_encoded_bytes(100, 2) == 32  # fle-test.cc:219,224
    assert L.fle_encoded_bytes(0, 5) == 40 and L.fle_encoded_bytes(64, 5) == 40 and L.fle_encoded_bytes(65, 5) == 80
    assert L.fle_strerror(-7).decode().startswith("no CUDA device")


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback(lib_path):
    import fle_b200
    with pytest.raises(fle_b200.FleError) as ei:
        fle_b200.FleContext(0)
    assert ei.value.code == fle_b200.capi.E_NODEVICE


def test_product_does_not_touch_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "impala-avx2-parquet-scanner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("oracle/", "oracle/") or "oracle" not in re.sub(
                    r"(#|//).*", "", txt).lower(), f"{f} mentions the oracle outside comments"
