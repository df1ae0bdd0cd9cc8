"""Helpers that drive the compiled reference (oracle/_ref) into the canonical arena layout.

TEST INFRASTRUCTURE ONLY.
"""
import zlib

import numpy as np

import oraclebind
import refbind


def crc(a) -> str:
    return "%08x" % (zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF)


def ref_plane_arena(o: oraclebind.Oracle, r: refbind.RefWavelet) -> np.ndarray:
    """Copy every band of a reference CWavelet2D into one arena laid out as the oracle/C-ABI does."""
    a = o.new_arena()
    for i in range(o.nbands):
        f, g = o.info(i), r.info(i)
        assert (f["dimx"], f["dimy"], f["stride"], f["is_int"]) == (g["dimx"], g["dimy"], g["stride"], g["is_int"])
        assert f["weight"] == g["weight"]
        o.band_view(a, i)[:, :f["dimx"]] = r.band(i)[:, :f["dimx"]]
    return a


def ref_encode_arenas(img_u8, q, levels=5, level_chg=None, trans=0):
    """colour + Transform + quantiser half of CodeBand for every plane, through the reference."""
    ch, h, w = img_u8.shape
    if level_chg is None:
        level_chg = levels - 4
    o = oraclebind.Oracle(w, h, levels, level_chg, trans=trans)
    planes = oraclebind.colour_fwd(img_u8, q)
    out = []
    for p in range(ch):
        r = refbind.RefWavelet(w, h, levels, level_chg, trans)
        r.transform(planes[p].copy())
        Q, lam = oraclebind.plane_quant(q, ch, p)
        r.quant(Q, lam)
        out.append(ref_plane_arena(o, r))
        r.close()
    return o, np.concatenate(out)


def ref_decode_arenas(o, arenas, ch, q, trans=0):
    """TSUQi + TransformI + inverse colour through the reference, from signed quantised arenas."""
    planes = np.zeros((ch, o.h, o.w), dtype=np.int16)
    for p in range(ch):
        r = refbind.RefWavelet(o.w, o.h, o.g.levels, o.g.level_chg, trans)
        a = arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes]
        for i in range(o.nbands):
            r.set_band(i, o.band_view(a, i))
        Q, _ = oraclebind.plane_quant(q, ch, p)
        if q:
            r.tsuqi(Q)
        planes[p] = r.transform_inv()
        r.close()
    return oraclebind.colour_inv(planes, q)


def ref_payload_from_arenas(o, arenas, ch, trans=0):
    """Run the reference's ENTROPY half of CodeBand (wavelet2d.cpp:119-159 minus the quantiser calls,
    oracle/ref_harness.cpp:entropy_half) over already-quantised band arenas, planes in ric's order
    (Y, Cg, Co -- ric.cpp:163-168), and return the payload exactly as ric writes it after the header."""
    import ctypes as C
    L = refbind.lib()
    buf = np.zeros(o.w * o.h * ch * 2 + 4096, dtype=np.uint8)
    codec = L.ref_codec_new_enc(buf.ctypes.data)
    order = [2, 1, 0] if ch == 3 else [0]
    for p in order:
        r = refbind.RefWavelet(o.w, o.h, o.g.levels, o.g.level_chg, trans)
        a = arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes]
        for i in range(o.nbands):
            r.set_band(i, o.band_view(a, i))
        L.ref_entropy_encode(r.h, codec)
        r.close()
    n = L.ref_codec_end(codec, buf.ctypes.data)
    L.ref_codec_free(codec)
    return buf[2:n].copy()
