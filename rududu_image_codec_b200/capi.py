"""ctypes binding of librududu_b200.so, the C-ABI library declared in include/ric_b200.h.

This is plumbing for the tests and bench.py: every compute call below ends in a hand-written sm_100a
kernel.  There is no fallback: if the shared library is missing, or no CUDA device is usable,
loading / `Context()` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librududu_b200.so")
MAX_LEVELS = 16
MAX_BANDS = 3 * MAX_LEVELS + 1

CDF97, CDF53, HAAR = 0, 1, 2
OK, E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4

EXPORTS = [
    "ric_create", "ric_destroy", "ric_get_info", "ric_get_band", "ric_last_error", "ric_quants",
    "ric_plane_quant", "ric_encode_u8", "ric_decode_u8", "ric_encode_u8_device", "ric_decode_u8_device",
    "ric_last_launch_count", "ric_transform", "ric_quant", "ric_tsuq", "ric_tsuqi", "ric_transform_inv",
    "ric_host_alloc", "ric_host_free", "ric_set_profiling", "ric_get_level_times", "ric_get_path_stats",
    "ric_encode_u8_stream", "ric_decode_u8_stream", "ric_sync", "ric_header_write", "ric_header_parse",
    "ric_entropy_encode", "ric_entropy_decode", "ric_compress_u8", "ric_decompress_u8",
    "ric_mux_encoder", "ric_mux_decoder", "ric_mux_code_plane", "ric_mux_decode_plane", "ric_mux_finish",
    "ric_mux_destroy", "ric_entropy_encode_device", "ric_entropy_decode_device",
    "ric_compress_u8_gpu", "ric_decompress_u8_gpu", "ric_entropy_encode_hinted",
    "ric_set_base_weight", "ric_quant_host", "ric_tsuq_host", "ric_buf_tsuq", "ric_buf_tsuqi", "ric_buf_build_tree",
]


class RicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ric_b200 error %d: %s" % (code, msg))
        self.code = code


class BandInfo(C.Structure):
    _fields_ = [("dimx", C.c_int), ("dimy", C.c_int), ("stride", C.c_int), ("is_int", C.c_int),
                ("weight", C.c_float), ("offset", C.c_size_t)]


class Info(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("levels", C.c_int),
                ("level_chg", C.c_int), ("align", C.c_int), ("trans", C.c_int), ("nlev", C.c_int),
                ("nbands", C.c_int), ("max_batch", C.c_int), ("arena_bytes", C.c_size_t),
                ("image_arena_bytes", C.c_size_t)]


CHUNK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)  # ric_chunk_fn

_lib = None


def lib():
    """Load the C-ABI library (built by __graft_entry__.build() / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.ric_create.argtypes = [C.POINTER(vp)] + [i] * 9
        L.ric_destroy.argtypes = [vp]
        L.ric_get_info.argtypes = [vp, C.POINTER(Info)]
        L.ric_get_band.argtypes = [vp, i, C.POINTER(BandInfo)]
        L.ric_last_error.restype = C.c_char_p
        L.ric_quants.argtypes = [i]
        L.ric_plane_quant.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
        L.ric_encode_u8.argtypes = [vp, vp, i, i, vp]
        L.ric_decode_u8.argtypes = [vp, vp, i, i, vp]
        L.ric_encode_u8_device.argtypes = [vp, vp, sz, i, i, vp, vp]
        L.ric_decode_u8_device.argtypes = [vp, vp, i, i, vp, sz, vp]
        L.ric_last_launch_count.argtypes = [vp]
        L.ric_transform.argtypes = [vp, vp, i, vp]
        L.ric_quant.argtypes = [vp, i, i, vp]
        L.ric_tsuq.argtypes = [vp, i, C.c_float, vp, C.POINTER(C.c_uint)]
        L.ric_tsuqi.argtypes = [vp, i, vp]
        L.ric_transform_inv.argtypes = [vp, vp, vp, i]
        L.ric_host_alloc.argtypes = [C.POINTER(vp), sz]
        L.ric_host_free.argtypes = [vp]
        L.ric_set_profiling.argtypes = [vp, i]
        L.ric_encode_u8_stream.argtypes = [vp, vp, i, i, vp, CHUNK_FN, vp]
        L.ric_decode_u8_stream.argtypes = [vp, vp, i, i, vp, CHUNK_FN, vp]
        L.ric_sync.argtypes = [vp]
        L.ric_header_write.argtypes = [vp, i, i, i, i, i]
        L.ric_header_parse.argtypes = [vp] + [C.POINTER(i)] * 5
        L.ric_get_level_times.argtypes = [vp, i, C.POINTER(C.c_float), i]
        L.ric_get_path_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), i]
        L.ric_entropy_encode.argtypes = [i] * 6 + [vp, vp, sz, C.POINTER(sz)]
        L.ric_entropy_decode.argtypes = [i] * 6 + [vp, sz, vp]
        L.ric_entropy_encode_hinted.argtypes = [i] * 6 + [vp, vp, sz, C.POINTER(sz)]
        L.ric_compress_u8.argtypes = [vp, vp, i, i, vp, sz, vp, i]
        L.ric_entropy_encode_device.argtypes = [vp, vp, i, vp, sz, vp, vp]
        L.ric_entropy_decode_device.argtypes = [vp, vp, sz, vp, i, vp, vp, vp]
        L.ric_compress_u8_gpu.argtypes = [vp, vp, i, i, vp, sz, vp]
        L.ric_decompress_u8_gpu.argtypes = [vp, vp, sz, vp, i, vp]
        L.ric_mux_encoder.argtypes = [C.POINTER(vp), vp, sz, C.c_uint]
        L.ric_mux_decoder.argtypes = [C.POINTER(vp), vp, sz]
        L.ric_mux_code_plane.argtypes = [vp] + [i] * 5 + [vp]
        L.ric_mux_decode_plane.argtypes = [vp] + [i] * 5 + [vp]
        L.ric_mux_finish.argtypes = [vp, C.POINTER(sz)]
        L.ric_mux_destroy.argtypes = [vp]
        L.ric_decompress_u8.argtypes = [vp, vp, sz, vp, i, vp, i]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RicError(rc, lib().ric_last_error().decode())


def quants(idx):
    return lib().ric_quants(idx)


def plane_quant(q, channels, plane):
    a, b = C.c_int(), C.c_int()
    _check(lib().ric_plane_quant(q, channels, plane, C.byref(a), C.byref(b)))
    return a.value, b.value


def header_write(width, height, q, color, trans):
    buf = (C.c_uint8 * 9)()
    _check(lib().ric_header_write(buf, width, height, q, color, trans))
    return bytes(buf)


def header_parse(data):
    buf = (C.c_uint8 * 9).from_buffer_copy(bytes(data[:9]))
    v = [C.c_int() for _ in range(5)]
    _check(lib().ric_header_parse(buf, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


def entropy_encode(width, height, channels, image_arena, levels=5, level_chg=None, align=32, cap=None, hinted=False):
    """Host entropy stage: one image's quantised band arenas -> .ric payload (bytes after the header).
    The arenas are consumed (markers cleared in place).  No GPU involved.  hinted=True: the pre-pass + hinted
    walker the device stage uses (arenas untouched)."""
    if level_chg is None:
        level_chg = max(levels - 4, 0)
    if cap is None:
        cap = 2 * width * height * channels + 4096
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    fn = lib().ric_entropy_encode_hinted if hinted else lib().ric_entropy_encode
    _check(fn(width, height, channels, levels, level_chg, align, _ptr(image_arena), _ptr(out), cap, C.byref(n)))
    return out[:n.value].copy()


def entropy_decode(width, height, channels, payload, image_arena, levels=5, level_chg=None, align=32):
    """Host entropy stage, inverse: .ric payload -> signed quantised band arenas (decode-stage input)."""
    if level_chg is None:
        level_chg = max(levels - 4, 0)
    payload = np.ascontiguousarray(np.frombuffer(bytes(payload), dtype=np.uint8))
    _check(lib().ric_entropy_decode(width, height, channels, levels, level_chg, align, _ptr(payload),
                                    payload.size, _ptr(image_arena)))
    return image_arena


class Mux:
    """The plane-at-a-time entropy coder object (the reference's CMuxCodec + CodeBand/DecodeBand halves).
    `stream`: a numpy u8 buffer laid out like the reference's (payload from offset 2)."""

    def __init__(self, stream, encode, first_word=0):
        self.L, self.h, self.stream = lib(), C.c_void_p(), stream
        if encode:
            _check(self.L.ric_mux_encoder(C.byref(self.h), _ptr(stream), stream.size, first_word))
        else:
            _check(self.L.ric_mux_decoder(C.byref(self.h), _ptr(stream), stream.size))

    def code_plane(self, width, height, plane_arena, levels=5, level_chg=None, align=32):
        lc = max(levels - 4, 0) if level_chg is None else level_chg
        _check(self.L.ric_mux_code_plane(self.h, width, height, levels, lc, align, _ptr(plane_arena)))

    def decode_plane(self, width, height, plane_arena, levels=5, level_chg=None, align=32):
        lc = max(levels - 4, 0) if level_chg is None else level_chg
        _check(self.L.ric_mux_decode_plane(self.h, width, height, levels, lc, align, _ptr(plane_arena)))

    def finish(self):
        n = C.c_size_t()
        _check(self.L.ric_mux_finish(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            self.L.ric_mux_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Context:
    """One (width, height, channels, levels, transform) configuration on one GPU: the device-side
    counterpart of a set of reference `CWavelet2D` objects (src/lib/wavelet2d.h:27-88)."""

    def __init__(self, width, height, channels=1, levels=5, level_chg=None, align=32, trans=CDF97,
                 max_batch=1, device=0):
        self.L = lib()
        if level_chg is None:
            level_chg = max(levels - 4, 0)  # ric.cpp:159
        h = C.c_void_p()
        _check(self.L.ric_create(C.byref(h), device, width, height, channels, levels, level_chg, align,
                                 trans, max_batch))
        self.h = h
        inf = Info()
        _check(self.L.ric_get_info(self.h, C.byref(inf)))
        self.info = inf
        self.width, self.height, self.channels = width, height, channels
        self.nlev, self.nbands = inf.nlev, inf.nbands
        self.arena_bytes, self.image_arena_bytes = inf.arena_bytes, inf.image_arena_bytes
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "h", None):
            self.L.ric_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def band(self, i):
        b = BandInfo()
        _check(self.L.ric_get_band(self.h, i, C.byref(b)))
        return dict(dimx=b.dimx, dimy=b.dimy, stride=b.stride, is_int=b.is_int, weight=b.weight,
                    offset=b.offset, size=4 if b.is_int else 2)

    def band_view(self, arena, i, plane=0, image=0):
        f = self.band(i)
        dt = np.int32 if f["is_int"] else np.int16
        o = image * self.image_arena_bytes + plane * self.arena_bytes + f["offset"]
        return arena[o:o + f["stride"] * f["dimy"] * f["size"]].view(dt).reshape(f["dimy"], f["stride"])

    # ---- whole-stage, host buffers ------------------------------------------------------------
    def encode_u8(self, imgs, q, out=None):
        """imgs: u8 (n, channels, height, width) -> arenas u8 (n * image_arena_bytes)."""
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8).reshape(-1, self.channels, self.height, self.width)
        n = imgs.shape[0]
        if out is None:
            out = np.empty(n * self.image_arena_bytes, dtype=np.uint8)
        _check(self.L.ric_encode_u8(self.h, _ptr(imgs), n, q, _ptr(out)))
        return out

    def decode_u8(self, arenas, n, q, out=None):
        """arenas: signed quantised coefficients (DecodeBand output) -> u8 (n, channels, height, width)."""
        if out is None:
            out = np.empty((n, self.channels, self.height, self.width), dtype=np.uint8)
        _check(self.L.ric_decode_u8(self.h, _ptr(arenas), n, q, _ptr(out)))
        return out

    def encode_u8_stream(self, imgs, q, out, done):
        """Asynchronous encode: `done(first_image, n_images)` fires per chunk once its arenas are in `out`
        (called on a CUDA callback thread).  Call sync() before reusing the buffers."""
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8).reshape(-1, self.channels, self.height, self.width)
        self._keep = (imgs, out, CHUNK_FN(lambda user, first, cnt: done(first, cnt)))
        _check(self.L.ric_encode_u8_stream(self.h, _ptr(imgs), imgs.shape[0], q, _ptr(out), self._keep[2], None))

    def decode_u8_stream(self, arenas, n, q, out, done):
        self._keep = (arenas, out, CHUNK_FN(lambda user, first, cnt: done(first, cnt)))
        _check(self.L.ric_decode_u8_stream(self.h, _ptr(arenas), n, q, _ptr(out), self._keep[2], None))

    def sync(self):
        _check(self.L.ric_sync(self.h))

    # ---- whole .ric files -----------------------------------------------------------------------
    def compress_u8(self, imgs, q, threads=0, stride=None, entropy_on_device=False):
        """imgs: u8 (n, channels, height, width) -> list of n complete .ric files (bytes)."""
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8).reshape(-1, self.channels, self.height, self.width)
        n = imgs.shape[0]
        if stride is None:
            stride = self.width * self.height * self.channels * 2 + 4096
        files = np.empty((n, stride), dtype=np.uint8)
        sizes = np.zeros(n, dtype=np.uint64)
        if entropy_on_device:
            _check(self.L.ric_compress_u8_gpu(self.h, _ptr(imgs), n, q, _ptr(files), stride, _ptr(sizes)))
        else:
            _check(self.L.ric_compress_u8(self.h, _ptr(imgs), n, q, _ptr(files), stride, _ptr(sizes), threads))
        return [files[i, :int(sizes[i])].tobytes() for i in range(n)]

    def decompress_u8(self, files, threads=0, entropy_on_device=False):
        """files: list of .ric files (bytes) of this context's geometry -> u8 (n, channels, height, width)."""
        n = len(files)
        stride = max(len(f) for f in files)
        buf = np.zeros((n, stride), dtype=np.uint8)
        sizes = np.zeros(n, dtype=np.uint64)
        for i, f in enumerate(files):
            buf[i, :len(f)] = np.frombuffer(f, dtype=np.uint8)
            sizes[i] = len(f)
        out = np.empty((n, self.channels, self.height, self.width), dtype=np.uint8)
        if entropy_on_device:
            _check(self.L.ric_decompress_u8_gpu(self.h, _ptr(buf), stride, _ptr(sizes), n, _ptr(out)))
        else:
            _check(self.L.ric_decompress_u8(self.h, _ptr(buf), stride, _ptr(sizes), n, _ptr(out), threads))
        return out

    # ---- device-resident variants (pointers are raw device addresses, stream a cudaStream_t) ----
    def encode_u8_device(self, d_src, pitch, n, q, d_arenas, stream=0):
        _check(self.L.ric_encode_u8_device(self.h, d_src, pitch, n, q, d_arenas, stream))

    def decode_u8_device(self, d_arenas, n, q, d_dst, pitch, stream=0):
        _check(self.L.ric_decode_u8_device(self.h, d_arenas, n, q, d_dst, pitch, stream))

    def entropy_encode_device(self, d_arenas, n, d_out, stride, d_sizes, stream=0):
        _check(self.L.ric_entropy_encode_device(self.h, _ptr(d_arenas), n, _ptr(d_out), stride, _ptr(d_sizes), stream))

    def entropy_decode_device(self, d_payloads, stride, d_sizes, n, d_arenas, stream=0, d_status=None):
        _check(self.L.ric_entropy_decode_device(self.h, _ptr(d_payloads), stride, _ptr(d_sizes), n, _ptr(d_arenas),
                                                _ptr(d_status) if d_status is not None else None, stream))

    def set_profiling(self, on=True):
        _check(self.L.ric_set_profiling(self.h, int(on)))

    def path_stats(self):
        """Packed-kernel path counters since the last call (profiling must be on): [fwd iterations, fwd scalar, inv iterations, inv scalar column passes, inv scalar row passes]."""
        buf = (C.c_ulonglong * 5)()
        _check(self.L.ric_get_path_stats(self.h, buf, 5))
        return [int(v) for v in buf]

    def level_times(self, direction):
        """Per-launch durations (ms) of the last encode (0) / decode (1) *_device call."""
        buf = (C.c_float * MAX_LEVELS)()
        n = self.L.ric_get_level_times(self.h, direction, buf, MAX_LEVELS)
        if n < 0:
            _check(n)
        return [buf[k] for k in range(n)]

    def last_launch_count(self):
        return self.L.ric_last_launch_count(self.h)

    # ---- plane-level calls mirroring the reference class API -------------------------------------
    def transform(self, plane):
        """CWavelet2D::Transform<short> on one int16 plane; returns the arena of raw coefficients."""
        p = np.ascontiguousarray(plane, dtype=np.int16)
        a = np.empty(self.arena_bytes, dtype=np.uint8)
        _check(self.L.ric_transform(self.h, _ptr(p), p.shape[1], _ptr(a)))
        return a

    def quant(self, Quant, lam):
        """Quantiser half of CodeBand on the bands left on the device by transform()."""
        a = np.empty(self.arena_bytes, dtype=np.uint8)
        _check(self.L.ric_quant(self.h, Quant, lam, _ptr(a)))
        return a

    def tsuq(self, Quant, thres):
        a = np.empty(self.arena_bytes, dtype=np.uint8)
        cnt = C.c_uint()
        _check(self.L.ric_tsuq(self.h, Quant, thres, _ptr(a), C.byref(cnt)))
        return a, cnt.value

    def tsuqi(self, arena, Quant):
        a = np.ascontiguousarray(arena).copy()
        _check(self.L.ric_tsuqi(self.h, Quant, _ptr(a)))
        return a

    def transform_inv(self, arena):
        out = np.empty((self.height, self.width), dtype=np.int16)
        a = np.ascontiguousarray(arena)
        _check(self.L.ric_transform_inv(self.h, _ptr(a), _ptr(out), self.width))
        return out
