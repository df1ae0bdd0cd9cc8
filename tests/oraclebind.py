"""ctypes binding of oracle/libric_oracle.so (our C restatement of the hot path).

TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke()
may import this; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.abspath(os.path.join(_HERE, "..", "oracle"))
ORACLE_SO = os.path.join(ORACLE_DIR, "libric_oracle.so")
MAX_LEVELS = 16
MAX_BANDS = 3 * MAX_LEVELS + 1


class Band(C.Structure):
    _fields_ = [("dimx", C.c_int), ("dimy", C.c_int), ("stride", C.c_int), ("is_int", C.c_int),
                ("weight", C.c_float), ("offset", C.c_size_t)]


class Geom(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("levels", C.c_int), ("level_chg", C.c_int),
                ("align", C.c_int), ("trans", C.c_int), ("nlev", C.c_int),
                ("lev_w", C.c_int * MAX_LEVELS), ("lev_h", C.c_int * MAX_LEVELS),
                ("lev_is_int", C.c_int * MAX_LEVELS), ("nbands", C.c_int),
                ("band", Band * MAX_BANDS), ("arena_bytes", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libric_oracle.so"])
        L = C.CDLL(ORACLE_SO)
        G = C.POINTER(Geom)
        L.rico_quants.argtypes = [C.c_int]
        L.rico_geom_init.argtypes = [G] + [C.c_int] * 6
        L.rico_forward.argtypes = [G, C.c_void_p, C.c_int, C.c_void_p]
        L.rico_inverse.argtypes = [G, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rico_quant.argtypes = [G, C.c_void_p, C.c_int, C.c_int]
        L.rico_tsuq_all.argtypes = [G, C.c_void_p, C.c_int, C.c_float]
        L.rico_tsuq_all.restype = C.c_uint
        L.rico_tsuqi.argtypes = [G, C.c_void_p, C.c_int]
        L.rico_unfold.argtypes = [G, C.c_void_p]
        L.rico_colour_fwd.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        L.rico_colour_inv.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        L.rico_encode_image.argtypes = [G, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.rico_decode_image.argtypes = [G, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.rico_plane_quant.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 2
        _lib = L
    return _lib


class Oracle:
    """Geometry + arena helpers around the C restatement."""

    def __init__(self, w, h, levels=5, level_chg=None, align=32, trans=0):
        self.L = lib()
        if level_chg is None:
            level_chg = levels - 4
        self.g = Geom()
        rc = self.L.rico_geom_init(C.byref(self.g), w, h, levels, level_chg, align, trans)
        if rc != 0:
            raise ValueError("bad geometry")
        self.w, self.h = w, h
        self.nbands = self.g.nbands
        self.nlev = self.g.nlev
        self.arena_bytes = self.g.arena_bytes

    def info(self, i):
        b = self.g.band[i]
        return dict(dimx=b.dimx, dimy=b.dimy, stride=b.stride, is_int=b.is_int, weight=b.weight,
                    offset=b.offset, size=4 if b.is_int else 2)

    def new_arena(self, n=1):
        return np.zeros(self.arena_bytes * n, dtype=np.uint8)

    def band_view(self, arena, i, plane=0):
        f = self.info(i)
        dt = np.int32 if f["is_int"] else np.int16
        o = plane * self.arena_bytes + f["offset"]
        return arena[o:o + f["stride"] * f["dimy"] * f["size"]].view(dt).reshape(f["dimy"], f["stride"])

    def forward(self, plane):
        p = np.ascontiguousarray(plane, dtype=np.int16)
        a = self.new_arena()
        self.L.rico_forward(C.byref(self.g), p.ctypes.data, p.shape[1], a.ctypes.data)
        return a

    def inverse(self, arena, q1_quirk=1):
        out = np.zeros((self.h, self.w), dtype=np.int16)
        self.L.rico_inverse(C.byref(self.g), arena.ctypes.data, out.ctypes.data, self.w, q1_quirk)
        return out

    def quant(self, arena, Quant, lam):
        self.L.rico_quant(C.byref(self.g), arena.ctypes.data, Quant, lam)

    def tsuqi(self, arena, Quant):
        self.L.rico_tsuqi(C.byref(self.g), arena.ctypes.data, Quant)

    def tsuq_all(self, arena, Quant, thres):
        return self.L.rico_tsuq_all(C.byref(self.g), arena.ctypes.data, Quant, thres)

    def unfold(self, arena):
        self.L.rico_unfold(C.byref(self.g), arena.ctypes.data)

    def encode_image(self, img_u8, q):
        ch = img_u8.shape[0]
        a = self.new_arena(ch)
        src = np.ascontiguousarray(img_u8)
        self.L.rico_encode_image(C.byref(self.g), src.ctypes.data, ch, q, a.ctypes.data)
        return a

    def decode_image(self, arenas, ch, q):
        dst = np.zeros((ch, self.h, self.w), dtype=np.uint8)
        self.L.rico_decode_image(C.byref(self.g), arenas.ctypes.data, ch, q, dst.ctypes.data)
        return dst


def plane_quant(q, ch, p):
    a, b = C.c_int(), C.c_int()
    lib().rico_plane_quant(q, ch, p, C.byref(a), C.byref(b))
    return a.value, b.value


def colour_fwd(img_u8, q):
    ch, h, w = img_u8.shape
    out = np.zeros((ch, h, w), dtype=np.int16)
    src = np.ascontiguousarray(img_u8)
    lib().rico_colour_fwd(src.ctypes.data, w, h, ch, q, out.ctypes.data)
    return out


def colour_inv(planes, q):
    ch, h, w = planes.shape
    out = np.zeros((ch, h, w), dtype=np.uint8)
    p = np.ascontiguousarray(planes, dtype=np.int16)
    lib().rico_colour_inv(p.ctypes.data, w, h, ch, q, out.ctypes.data)
    return out
