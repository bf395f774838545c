"""CPU-only: libsvsk.so loads without a GPU and exports every symbol include/svsk.h declares."""
import ctypes
import os
import re

import pytest

from ensemble_svs_with_interactions_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "svsk.h")).read()
    return sorted(set(re.findall(r"SVSK_API\s+[\w\s\*]*?\b(svsk_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 20
    dll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()
    _lib.lib()  # resolves every symbol with its argtypes


def test_version_and_error_string_without_gpu():
    l = _lib.lib()
    assert l.svsk_version() == 100
    import torch
    if not torch.cuda.is_available():
        rc = l.svsk_device_check(0)
        assert rc != 0  # no driver: a cudaError code, not a crash
        assert len(l.svsk_last_error()) > 0


def test_argument_errors_are_reported_before_launch():
    l = _lib.lib()
    assert l.svsk_conv1d_f32(None, None) == -1
    assert b"null" in l.svsk_last_error()
    assert l.svsk_diffnet_packed_row(0, 100) == -1          # C must be a multiple of 128
    assert l.svsk_diffnet_packed_row(5, 256) == 5           # gate rows of block 0 stay
    assert l.svsk_diffnet_packed_row(256, 256) == 128       # filter row 0 -> packed 128
    assert l.svsk_diffnet_packed_row(128, 256) == 256       # gate row 128 -> second pair
    assert l.svsk_diffnet_packed_row(256 + 128, 256) == 384


def test_no_cpu_fallback():
    import torch
    from ensemble_svs_with_interactions_b200.diffsinger import DiffNet
    m = DiffNet(in_dim=8, encoder_hidden_dim=8, residual_layers=2, residual_channels=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 8, 4), torch.zeros(1, dtype=torch.long), torch.zeros(1, 8, 4))
