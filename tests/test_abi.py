"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/mmg_b200.h declares
and validates its arguments (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge
from multimodalgame_b200 import capi, engine as eng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return capi.Library(capi.LIB_PATH)


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "mmg_b200.h")).read()
    declared = set(re.findall(r"\b(mmg_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib.dll, name), "symbol %s declared in include/mmg_b200.h but not exported" % name
    assert set(capi.Library.SYMBOLS) <= declared
    assert lib.dll.mmg_abi_version() == capi.MMG_ABI_VERSION == 6


def test_layouts_and_validation(lib):
    cfg = eng.make_config(batch=64, n_classes=30, img_feat_dim=2048, img_h_dim=256, baseline_hid_dim=500,
                          sender_out_dim=32, rec_hidden=64, rec_w_dim=32, wv_dim=100, max_exchange=10)
    L = capi.ParamLayout()
    lib.call("mmg_param_layout_get", C.byref(cfg), C.byref(L))
    # parameter counts of SURVEY.md §8a: sender 541 248, receiver 42 146, baselines 145 001 + 49 001
    n = lambda a, b: sum(L.rows[i] * L.cols[i] for i in range(a, b))
    assert n(0, 21) == 42146 and n(21, 29) == 541248 and n(29, 33) == 49001 and n(33, 37) == 145001
    assert all(L.offset[i] % 4 == 0 for i in range(capi.MMG_P_COUNT))
    assert [capi.PARAM_NAMES[i][0] for i in range(capi.MMG_P_COUNT)] == [capi.SEGMENTS[L.segment[i]] for i in range(capi.MMG_P_COUNT)]
    W = capi.WorkspaceLayout()
    lib.call("mmg_workspace_layout_get", C.byref(cfg), C.byref(W))
    assert W.total_bytes > 0 and W.stats % 8 == 0
    bad = eng.make_config(batch=0, n_classes=30)
    with pytest.raises(capi.MmgError):
        lib.call("mmg_param_layout_get", C.byref(bad), C.byref(L))
    with pytest.raises(AssertionError):       # model.py:1756
        eng.make_config(batch=4, n_classes=3, sender_out_dim=8, rec_w_dim=16)


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    capi._LIB = None
    with pytest.raises(capi.MmgError):
        capi.load()
