"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol that
include/ric_b200.h declares; host-side scalar helpers agree with the oracle; and the product fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

import oraclebind
from rududu_image_codec_b200 import capi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    src = open(os.path.join(ROOT, "include", "ric_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ric_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.EXPORTS) == names


def test_quants_table_matches_oracle():
    for idx in range(0, 64):
        assert capi.quants(idx) == oraclebind.lib().rico_quants(idx)
    # Quants(q+20), q = 1..31
    want = [32, 36, 42, 48, 56, 64, 72, 84, 96, 112, 128, 144, 168, 192, 224, 256, 288, 336, 384, 448, 512, 576,
            672, 768, 896, 1024, 1152, 1344, 1536, 1792, 2048]
    assert [capi.quants(q + 20) for q in range(1, 32)] == want
    for q in (0, 1, 9, 31):
        for ch in (1, 3):
            for p in range(ch):
                assert capi.plane_quant(q, ch, p) == oraclebind.plane_quant(q, ch, p)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(capi.RicError) as e:
        capi.Context(512, 512)
    assert e.value.code == capi.E_CUDA
    assert b"no CPU path" in capi.lib().ric_last_error() or b"CUDA" in capi.lib().ric_last_error()


def test_argument_validation_needs_no_gpu():
    h = ctypes.c_void_p()
    L = capi.lib()
    assert L.ric_create(ctypes.byref(h), 0, 8, 8, 1, 5, 1, 32, 0, 1) == capi.E_ARG
    assert L.ric_create(ctypes.byref(h), 0, 64, 64, 2, 5, 1, 32, 0, 1) == capi.E_ARG
    assert L.ric_create(ctypes.byref(h), 0, 64, 64, 1, 5, 5, 32, 0, 1) == capi.E_ARG
    assert L.ric_create(ctypes.byref(h), 0, 66, 64, 1, 5, 1, 32, 2, 1) == capi.E_UNSUPPORTED  # Haar, odd level width
    assert L.ric_create(ctypes.byref(h), 0, 64, 64, 1, 5, 1, 32, 3, 1) == capi.E_UNSUPPORTED  # unknown transform
    assert L.ric_destroy(None) == 0


def test_ric_container_header():
    """"RUD2", u16 LE width, u16 LE height, Quant:5 | Color<<5 | Trans<<6 (src/ric/ric.cpp:114-121,150-154)."""
    assert capi.header_write(3840, 2160, 9, 1, 0) == b"RUD2" + bytes([0x00, 0x0F, 0x70, 0x08, 9 | 1 << 5])
    assert capi.header_write(512, 512, 0, 0, 1) == b"RUD2" + bytes([0x00, 0x02, 0x00, 0x02, 1 << 6])
    for args in [(65535, 1, 31, 1, 2), (16, 65535, 0, 0, 0), (1920, 1080, 20, 1, 1)]:
        assert capi.header_parse(capi.header_write(*args)) == args
    with pytest.raises(capi.RicError):
        capi.header_parse(b"RUD1" + bytes(5))
    with pytest.raises(capi.RicError):
        capi.header_write(70000, 10, 9, 1, 0)
