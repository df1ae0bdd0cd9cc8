"""ctypes binding of oracle/_ref/libric_ref.so (the compiled, unmodified reference + our harness).

TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline/--impl reference leg and
__graft_entry__.smoke() may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "..", "oracle", "_ref", "libric_ref.so")


def available() -> bool:
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ref_wavelet_new.restype = C.c_void_p
        L.ref_wavelet_new.argtypes = [C.c_int] * 5
        L.ref_wavelet_free.argtypes = [C.c_void_p]
        L.ref_levels.argtypes = [C.c_void_p]
        L.ref_band_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                    C.POINTER(C.c_void_p)]
        L.ref_band_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_band_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_transform_inv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_quant.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_tsuqi.argtypes = [C.c_void_p, C.c_int]
        L.ref_tsuq.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_tsuq.restype = C.c_uint
        L.ref_codec_new_enc.restype = C.c_void_p
        L.ref_codec_new_enc.argtypes = [C.c_void_p]
        L.ref_codec_new_dec.restype = C.c_void_p
        L.ref_codec_new_dec.argtypes = [C.c_void_p]
        L.ref_codec_end.restype = C.c_long
        L.ref_codec_end.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_codec_free.argtypes = [C.c_void_p]
        L.ref_codeband.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_entropy_encode.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_decodeband.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_compress.restype = C.c_long
        L.ref_compress.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_long]
        L.ref_decompress.argtypes = [C.c_void_p, C.c_long] + [C.c_int] * 7 + [C.c_void_p]
        L.ref_bench_stage.restype = C.c_double
        L.ref_bench_stage.argtypes = [C.c_void_p] + [C.c_int] * 10
        L.ref_bench_stage_split.argtypes = [C.POINTER(C.c_double)]
        L.ref_bench_codec.restype = C.c_double
        L.ref_bench_codec.argtypes = [C.c_void_p] + [C.c_int] * 10
        _lib = L
    return _lib


class RefWavelet:
    """One reference CWavelet2D (+SetWeight) with numpy access to its bands (canonical order)."""

    def __init__(self, w, h, levels, level_chg, trans=0):
        self.L = lib()
        self.h = self.L.ref_wavelet_new(w, h, levels, level_chg, trans)
        self.w, self.hgt = w, h
        self.nlev = self.L.ref_levels(self.h)
        self.nbands = 3 * self.nlev + 1

    def close(self):
        if self.h:
            self.L.ref_wavelet_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def info(self, i):
        a = (C.c_int * 5)()
        wt = C.c_float()
        self.L.ref_band_info(self.h, i, a, C.byref(wt), None)
        return dict(dimx=a[0], dimy=a[1], stride=a[2], is_int=a[3], size=a[4], weight=wt.value)

    def band(self, i):
        f = self.info(i)
        arr = np.zeros((f["dimy"], f["stride"]), dtype=np.int32 if f["is_int"] else np.int16)
        self.L.ref_band_get(self.h, i, arr.ctypes.data)
        return arr

    def set_band(self, i, arr):
        f = self.info(i)
        a = np.ascontiguousarray(arr, dtype=np.int32 if f["is_int"] else np.int16)
        assert a.shape == (f["dimy"], f["stride"])
        self.L.ref_band_set(self.h, i, a.ctypes.data)

    def transform(self, plane):
        """plane: int16 (H, W) array; destroyed. Needs slack after the plane for the int scratch."""
        buf = np.zeros(plane.size + 64, dtype=np.int16)
        buf[:plane.size] = plane.reshape(-1)
        self.L.ref_transform(self.h, buf.ctypes.data, plane.shape[1])

    def transform_inv(self):
        buf = np.zeros(self.w * self.hgt + 64, dtype=np.int16)
        self.L.ref_transform_inv(self.h, buf.ctypes.data, self.w, self.hgt)
        return buf[:self.w * self.hgt].reshape(self.hgt, self.w).copy()

    def quant(self, Quant, lam):
        self.L.ref_quant(self.h, Quant, lam)

    def tsuqi(self, Quant):
        self.L.ref_tsuqi(self.h, Quant)


def compress(img_u8, q, trans=0, levels=5, level_chg=None):
    c, h, w = img_u8.shape
    if level_chg is None:
        level_chg = levels - 4
    out = np.zeros(img_u8.size * 2 + 4096, dtype=np.uint8)
    src = np.ascontiguousarray(img_u8)
    n = lib().ref_compress(src.ctypes.data, w, h, c, q, trans, levels, level_chg, out.ctypes.data, out.size)
    assert n >= 0
    return out[:n].copy()


def decompress(payload, w, h, c, q, trans=0, levels=5, level_chg=None):
    if level_chg is None:
        level_chg = levels - 4
    dst = np.zeros((c, h, w), dtype=np.uint8)
    p = np.ascontiguousarray(payload)
    lib().ref_decompress(p.ctypes.data, p.size, w, h, c, q, trans, levels, level_chg, dst.ctypes.data)
    return dst
