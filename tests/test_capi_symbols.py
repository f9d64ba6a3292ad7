"""The C-ABI shared library loads on a CPU-only box and exports every symbol declared in
include/saicv_b200.h; the ctypes table in _lib.py lists the same set (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'saicv_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(saicv_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from simpleaicv_pytorch_training_examples_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert lib.saicv_version() >= 100


def test_ctypes_table_matches_header():
    from simpleaicv_pytorch_training_examples_b200 import _lib
    declared = set(_declared())
    bound = set(_lib.SIGNATURES) | {'saicv_last_error', 'saicv_launch_count'}
    assert declared == bound, (declared - bound, bound - declared)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from simpleaicv_pytorch_training_examples_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    import pytest
    with pytest.raises(RuntimeError, match='no CPU or library fallback'):
        _lib.load()
