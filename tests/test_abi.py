"""CPU-only: the C-ABI library loads and exports every symbol include/mrfp_b200.h declares."""
import os
import re

from tests.common import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mrfp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mrfp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mrfp_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert "mrfp_npplus_fwd_f32" in names and "mrfp_hrfp_fwd" in names
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))


def test_strerror_and_version_without_gpu():
    from mrfp_b200 import _lib
    lib = _lib.load()
    assert lib.mrfp_version() >= 100
    assert lib.mrfp_strerror(0) == b"success"
    assert b"shape" in lib.mrfp_strerror(-2)
    assert lib.mrfp_npplus_ws_bytes(8, 256, 36864) >= 8 * 256 * 5 * 8
    assert lib.mrfp_npplus_ws_bytes(0, 1, 1) == 0
