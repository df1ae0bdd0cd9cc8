"""Pins oracle/ric_oracle.c (our C restatement) against the compiled, unmodified reference
(oracle/_ref/libric_ref.so: /root/reference/src/lib/*.cpp + oracle/ref_harness.cpp).

CPU only.  The reference ships no tests, so these differential runs plus the
known-answer values of SURVEY.md Appendix C (tests/golden/kats.json) are what pin parity.
"""
import numpy as np
import pytest

import oraclebind
import refbind
from refutil import crc, ref_decode_arenas, ref_encode_arenas, ref_payload_from_arenas, ref_plane_arena
from rududu_image_codec_b200.synth import synth_image

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libric_ref.so not built")

SHAPES = [(512, 512), (517, 389), (64, 48), (33, 47), (250, 131), (16, 16), (129, 130), (96, 17)]


def _bands_equal(o, a, b):
    for i in range(o.nbands):
        f = o.info(i)
        x, y = o.band_view(a, i)[:, :f["dimx"]], o.band_view(b, i)[:, :f["dimx"]]
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)[0]
            return "band %d differs first at %s: %d vs %d" % (i, bad, x[tuple(bad)], y[tuple(bad)])
    return None


@pytest.mark.parametrize("w,h", SHAPES + [(1921, 1081), (3840, 2160)])
@pytest.mark.parametrize("levels,chg", [(5, 1), (5, 0), (6, 2), (3, 2)])
def test_geometry_and_weights(w, h, levels, chg):
    for trans in (0, 1):
        o = oraclebind.Oracle(w, h, levels, chg, trans=trans)
        r = refbind.RefWavelet(w, h, levels, chg, trans)
        assert o.nlev == r.nlev
        for i in range(o.nbands):
            f, g = o.info(i), r.info(i)
            for k in ("dimx", "dimy", "stride", "is_int", "weight"):
                assert f[k] == g[k], (i, k)
        r.close()


@pytest.mark.parametrize("w,h", SHAPES)
@pytest.mark.parametrize("levels,chg,trans", [(5, 1, 0), (5, 0, 0), (6, 2, 0), (4, 3, 0), (5, 1, 1), (5, 0, 1)])
def test_forward_full_range(w, h, levels, chg, trans):
    rng = np.random.default_rng(w * 7919 + h + levels + 10 * trans)
    plane = rng.integers(-32768, 32768, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, levels, chg, trans=trans)
    r = refbind.RefWavelet(w, h, levels, chg, trans)
    r.transform(plane.copy())
    assert _bands_equal(o, o.forward(plane), ref_plane_arena(o, r)) is None
    r.close()


@pytest.mark.parametrize("w,h", SHAPES)
@pytest.mark.parametrize("levels,chg,trans", [(5, 1, 0), (5, 0, 0), (6, 2, 0), (5, 1, 1)])
def test_inverse_full_range(w, h, levels, chg, trans):
    rng = np.random.default_rng(w * 31 + h + levels)
    o = oraclebind.Oracle(w, h, levels, chg, trans=trans)
    r = refbind.RefWavelet(w, h, levels, chg, trans)
    a = o.new_arena()
    for i in range(o.nbands):
        f = o.info(i)
        lo, hi = (-2 ** 20, 2 ** 20) if f["is_int"] else (-32768, 32768)
        v = o.band_view(a, i)
        v[:, :f["dimx"]] = rng.integers(lo, hi, size=(f["dimy"], f["dimx"]))
        r.set_band(i, v)
    assert np.array_equal(o.inverse(a, q1_quirk=1), r.transform_inv())
    r.close()


@pytest.mark.parametrize("w,h", [(512, 512), (517, 389), (250, 131), (33, 47), (129, 130)])
@pytest.mark.parametrize("q", [0, 1, 4, 9, 16, 31])
@pytest.mark.parametrize("ch", [1, 3])
def test_encode_quantiser(w, h, q, ch):
    img = synth_image(3, w, h, ch)
    trans = 1 if q == 0 else 0
    o, ref = ref_encode_arenas(img, q, trans=trans)
    ours = oraclebind.Oracle(w, h, 5, trans=trans).encode_image(img, q)
    for p in range(ch):
        s = slice(p * o.arena_bytes, (p + 1) * o.arena_bytes)
        assert _bands_equal(o, ours[s], ref[s]) is None, "plane %d" % p


@pytest.mark.parametrize("Quant,lam", [(7, 3), (96, 36), (1000, 400), (6144, 3000), (30000, 12000)])
def test_quantiser_random_coefficients(Quant, lam):
    """Heavy-tailed random planes: exercises ties in the candidate sort, every rank threshold,
    partial edge blocks, int32 coarse levels and (at large Quant) the int16 wrap paths."""
    w, h = 203, 151
    rng = np.random.default_rng(Quant)
    plane = (rng.standard_cauchy(size=(h, w)) * Quant / 6).clip(-32768, 32767).astype(np.int16)
    o = oraclebind.Oracle(w, h, 5, 1)
    r = refbind.RefWavelet(w, h, 5, 1, 0)
    r.transform(plane.copy())
    a = o.forward(plane)
    r.quant(Quant, lam)
    o.quant(a, Quant, lam)
    assert _bands_equal(o, a, ref_plane_arena(o, r)) is None
    r.close()


@pytest.mark.parametrize("w,h,ch", [(512, 512, 1), (320, 200, 3), (250, 132, 3)])
@pytest.mark.parametrize("q", [1, 9, 20, 31])
def test_decode_stage_matches_reference_decoder(w, h, ch, q):
    """unfold(encode bands) is what DecodeBand leaves in the arena: our TSUQi + TransformI +
    colour must give the pixels the reference decoder gives for the reference's own bitstream."""
    img = synth_image(1, w, h, ch)
    payload = refbind.compress(img, q)
    want = refbind.decompress(payload, w, h, ch, q)
    o = oraclebind.Oracle(w, h, 5)
    arenas = o.encode_image(img, q)
    for p in range(ch):
        o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
    assert np.array_equal(o.decode_image(arenas, ch, q), want)
    assert np.array_equal(ref_decode_arenas(o, arenas, ch, q), want)


@pytest.mark.parametrize("w,h,ch", [(512, 512, 1), (256, 192, 3)])
def test_lossless_round_trip(w, h, ch):
    img = synth_image(2, w, h, ch)
    o = oraclebind.Oracle(w, h, 5, trans=1)
    arenas = o.encode_image(img, 0)
    for p in range(ch):
        o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
    assert np.array_equal(o.decode_image(arenas, ch, 0), img)


def test_tsuq_all_and_tsuqi():
    w, h = 250, 131
    rng = np.random.default_rng(5)
    plane = rng.integers(-3000, 3000, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, 5, 1)
    r = refbind.RefWavelet(w, h, 5, 1, 0)
    r.transform(plane.copy())
    a = o.forward(plane)
    assert o.tsuq_all(a, 96, 0.7) == r.L.ref_tsuq(r.h, 96, 0.7)
    assert _bands_equal(o, a, ref_plane_arena(o, r)) is None
    r.tsuqi(96)
    o.tsuqi(a, 96)
    assert _bands_equal(o, a, ref_plane_arena(o, r)) is None
    r.close()


def test_kat_small(golden):
    """the compiled reference reproduces the survey's known answers."""
    for k in golden["kats"]:
        if k["w"] * k["h"] * k["ch"] > 1 << 21:
            continue
        img = synth_image(k["idx"], k["w"], k["h"], k["ch"])
        assert crc(img) == k["src_crc"]
        p = refbind.compress(img, k["q"], k["trans"], k["levels"])
        assert (len(p), crc(p)) == (k["payload_bytes"], k["payload_crc"])
        d = refbind.decompress(p, k["w"], k["h"], k["ch"], k["q"], k["trans"], k["levels"])
        assert crc(d) == k["dec8_crc"]


@pytest.mark.parametrize("w,h,ch,q", [(512, 512, 1, 9), (320, 200, 3, 9), (246, 131, 3, 4), (512, 384, 3, 27)])
def test_quant_half_plus_entropy_half_is_codeband(w, h, ch, q):
    """The split INTEGRATION.md relies on: quantiser half (GPU-able) then entropy half (host) gives
    the byte-identical payload of the reference's monolithic CodeBand -- also when the quantised
    arenas come from our restatement instead of the reference's buildTree."""
    img = synth_image(5, w, h, ch)
    want = refbind.compress(img, q)
    o, arenas = ref_encode_arenas(img, q)
    assert np.array_equal(ref_payload_from_arenas(o, arenas, ch), want)
    ours = oraclebind.Oracle(w, h, 5).encode_image(img, q)
    assert np.array_equal(ref_payload_from_arenas(o, ours, ch), want)


@pytest.mark.parametrize("w,h,levels,chg", [(512, 512, 5, 1), (320, 192, 5, 1), (64, 96, 3, 0), (640, 352, 5, 0)])
def test_haar_even_sizes(w, h, levels, chg):
    """Haar (src/lib/wavelet2d.cpp:766-855) where it is well defined: every level has even width and
    height (odd trailing rows/columns are left unprocessed by the reference)."""
    rng = np.random.default_rng(w + h)
    plane = rng.integers(-32768, 32768, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, levels, chg, trans=2)
    r = refbind.RefWavelet(w, h, levels, chg, 2)
    r.transform(plane.copy())
    a = o.forward(plane)
    assert _bands_equal(o, a, ref_plane_arena(o, r)) is None
    for i in range(o.nbands):
        r.set_band(i, o.band_view(a, i))
    assert np.array_equal(o.inverse(a), r.transform_inv())
    assert np.array_equal(o.inverse(a), plane)
    r.close()
