"""Host entropy stage of the .ric format against the compiled reference.

No GPU involved on either side: the quantised band arenas come from the reference's transform +
quantiser, then (a) our ric_entropy_encode must produce the reference's payload byte for byte
(CompressImage's stream, ric.cpp:123-178) and (b) our ric_entropy_decode must rebuild, from that
payload, exactly the band contents the reference's DecodeBand leaves (wavelet2d.cpp:183-221)."""
import ctypes as C

import numpy as np
import pytest

import oraclebind
import refbind
import refutil
from rududu_image_codec_b200 import capi, synth

needs_ref = pytest.mark.skipif(not refbind.available(), reason="compiled reference not built")


def _image(w, h, ch, idx=0):
    return synth.synth_image(idx, w, h, ch)


def _ref_decode_arenas(o, payload, ch, trans=0):
    """DecodeBand for every plane through the reference; returns arenas in the canonical layout."""
    L = refbind.lib()
    buf = np.zeros(len(payload) + o.w * o.h * ch * 2 + 4096, dtype=np.uint8)
    buf[2:2 + len(payload)] = np.frombuffer(payload, dtype=np.uint8)
    codec = L.ref_codec_new_dec(buf.ctypes.data)
    out = np.zeros(ch * o.arena_bytes, dtype=np.uint8)
    for p in ([2, 1, 0] if ch == 3 else [0]):
        r = refbind.RefWavelet(o.w, o.h, o.g.levels, o.g.level_chg, trans)
        L.ref_decodeband(r.h, codec)
        out[p * o.arena_bytes:(p + 1) * o.arena_bytes] = refutil.ref_plane_arena(o, r)
        r.close()
    L.ref_codec_free(codec)
    return out


CASES = [
    # w, h, ch, q, levels, level_chg, trans
    (512, 512, 1, 9, 5, 1, 0),
    (256, 192, 3, 9, 5, 1, 0),
    (246, 131, 3, 5, 5, 1, 0),      # partial blocks right and bottom at several levels
    (250, 134, 1, 13, 5, 1, 0),
    (333, 217, 3, 1, 5, 1, 0),
    (128, 96, 3, 31, 5, 1, 0),      # nearly everything insignificant
    (200, 120, 1, 0, 5, 1, 1),      # lossless 5/3
    (200, 120, 3, 0, 5, 1, 1),
    (320, 240, 1, 9, 6, 2, 0),      # two int levels
    (320, 240, 1, 9, 3, 0, 0),      # all-short
    (64, 64, 1, 20, 1, 0, 0),       # a single level: the finest band has no parent
    (192, 128, 3, 9, 5, 1, 2),      # Haar
    (1021, 67, 1, 7, 5, 1, 0),
]


@needs_ref
@pytest.mark.parametrize("w,h,ch,q,levels,level_chg,trans", CASES)
def test_payload_and_decoded_bands_match_reference(w, h, ch, q, levels, level_chg, trans):
    img = _image(w, h, ch)
    want = refbind.compress(img, q, trans=trans, levels=levels, level_chg=level_chg)
    o, arenas = refutil.ref_encode_arenas(img, q, levels=levels, level_chg=level_chg, trans=trans)
    got = capi.entropy_encode(w, h, ch, arenas.copy(), levels=levels, level_chg=level_chg)
    assert got.tobytes() == bytes(want)
    try:
        ref_dec = _ref_decode_arenas(o, bytes(want), ch, trans)
    except Exception:
        pytest.skip("reference decoder failed on its own stream")
    mine = np.full(ch * o.arena_bytes, 0xA5, dtype=np.uint8)
    capi.entropy_decode(w, h, ch, bytes(want), mine, levels=levels, level_chg=level_chg)
    for p in range(ch):
        for i in range(o.nbands):
            a = o.band_view(mine[p * o.arena_bytes:(p + 1) * o.arena_bytes], i)
            b = o.band_view(ref_dec[p * o.arena_bytes:(p + 1) * o.arena_bytes], i)
            f = o.info(i)
            assert np.array_equal(a[:, :f["dimx"]], b[:, :f["dimx"]]), (p, i)


def test_encoder_reports_small_buffer():
    img = _image(128, 128, 1)
    arenas = oraclebind.Oracle(128, 128, 5).encode_image(img, 5)
    with pytest.raises(capi.RicError) as e:
        capi.entropy_encode(128, 128, 1, arenas.copy(), cap=64)
    assert e.value.code == capi.E_NOMEM


@pytest.mark.parametrize("hinted", [False, True])
def test_encoder_rejects_arena_without_block_marks(hinted):
    """An arena no encode stage produces (all-zero blocks that lack the INSIGNIF_BLOCK mark, e.g. a staging buffer
    read before its copy landed) makes the coder report failure instead of indexing its count tables at -1."""
    o = oraclebind.Oracle(128, 128, 5)
    with pytest.raises(capi.RicError):
        capi.entropy_encode(128, 128, 1, np.zeros(o.arena_bytes, dtype=np.uint8), hinted=hinted)


def test_decoder_rejects_truncated_payload():
    img = _image(128, 128, 1)
    o = oraclebind.Oracle(128, 128, 5)
    want = capi.entropy_encode(128, 128, 1, o.encode_image(img, 5)).tobytes()
    with pytest.raises(capi.RicError):
        capi.entropy_decode(128, 128, 1, want[:len(want) // 4], np.zeros(o.arena_bytes, dtype=np.uint8))


def test_golden_payloads_from_oracle_arenas(golden):
    """Known answers of SURVEY Appendix C (payload size + CRC32, signed-arena CRC32; tests/golden/kats.json):
    band arenas from the oracle's transform+quantiser, payload from the product's entropy stage.  Runs
    without the compiled reference, so it also pins the entropy stage on the GPU box."""
    from refutil import crc
    done = 0
    for k in golden["kats"]:
        w, h, ch, q = k["w"], k["h"], k["ch"], k["q"]
        img = synth.synth_image(k["idx"], w, h, ch)
        o = oraclebind.Oracle(w, h, k["levels"], trans=k["trans"])
        arenas = o.encode_image(img, q)
        assert crc(arenas) == k["enc_arena_crc"]
        payload = capi.entropy_encode(w, h, ch, arenas.copy(), levels=k["levels"])
        assert (len(payload), crc(payload)) == (k["payload_bytes"], k["payload_crc"]), k
        assert capi.entropy_encode(w, h, ch, arenas, levels=k["levels"], hinted=True).tobytes() == payload.tobytes(), k
        back = np.full(arenas.size, 0x3C, dtype=np.uint8)
        capi.entropy_decode(w, h, ch, payload, back, levels=k["levels"])
        assert crc(back) == k["dec_arena_crc"], k
        done += 1
    assert done == len(golden["kats"])


@needs_ref
@pytest.mark.parametrize("ch,first_word", [(3, 0), (1, 0), (1, 0x1234)])
def test_plane_at_a_time_coder_matches_reference_buffer(ch, first_word):
    """ric_mux_*: the CMuxCodec-shaped object.  The whole stream buffer -- including bytes 0-1, which carry
    the coder's start word and never reach a .ric file -- equals the reference's, plane by plane."""
    w, h, q = 200, 136, 7
    img = _image(w, h, ch, 4)
    o, arenas = refutil.ref_encode_arenas(img, q)
    L = refbind.lib()
    L.ref_codec_new_enc_word.restype = C.c_void_p
    L.ref_codec_new_enc_word.argtypes = [C.c_void_p, C.c_int]
    ref_buf = np.zeros(w * h * ch * 2 + 4096, dtype=np.uint8)
    codec = L.ref_codec_new_enc_word(ref_buf.ctypes.data, first_word)
    mine_buf = np.zeros_like(ref_buf)
    mux = capi.Mux(mine_buf, encode=True, first_word=first_word)
    for p in ([2, 1, 0] if ch == 3 else [0]):
        a = arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes]
        r = refbind.RefWavelet(w, h, 5, 1, 0)
        for i in range(o.nbands):
            r.set_band(i, o.band_view(a, i))
        L.ref_entropy_encode(r.h, codec)
        r.close()
        mux.code_plane(w, h, a.copy())
    n = L.ref_codec_end(codec, ref_buf.ctypes.data)
    L.ref_codec_free(codec)
    assert mux.finish() == n
    mux.close()
    assert np.array_equal(mine_buf[:n], ref_buf[:n])
    # and back, plane by plane
    dec = capi.Mux(mine_buf[:n + 8], encode=False)
    ref_dec = _ref_decode_arenas(o, bytes(ref_buf[2:n]), ch)
    for p in ([2, 1, 0] if ch == 3 else [0]):
        a = np.full(o.arena_bytes, 0x77, dtype=np.uint8)
        dec.decode_plane(w, h, a)
        for i in range(o.nbands):
            f = o.info(i)
            assert np.array_equal(o.band_view(a, i)[:, :f["dimx"]],
                                  o.band_view(ref_dec[p * o.arena_bytes:(p + 1) * o.arena_bytes], i)[:, :f["dimx"]])
    dec.close()


def test_mux_argument_checks():
    buf = np.zeros(64, dtype=np.uint8)
    enc = capi.Mux(buf, encode=True)
    with pytest.raises(capi.RicError):
        enc.decode_plane(64, 64, np.zeros(16, dtype=np.uint8))   # an encoder cannot decode
    with pytest.raises(capi.RicError):
        enc.code_plane(8, 8, np.zeros(16, dtype=np.uint8))       # geometry below the codec's minimum
    enc.close()
    with pytest.raises(capi.RicError):
        capi.Mux(np.zeros(4, dtype=np.uint8), encode=True)       # no room for a stream


@needs_ref
def test_random_geometries_match_reference():
    """Seeded sweep over sizes, level counts, level_chg, q, transform and image statistics (texture, white noise,
    two-tone): payload byte-identical to the reference, and our decoder inverts our encoder."""
    import random
    rnd = random.Random(7)
    done = 0
    while done < 24:
        levels = rnd.choice([1, 2, 3, 4, 5, 5, 6])
        w, h, ch = rnd.randint(16, 200), rnd.randint(16, 160), rnd.choice([1, 3])
        q = rnd.choice([0, 1, 5, 9, 13, 27, 31])
        trans = rnd.choice([0, 0, 1]) if q else 1
        lc = rnd.randint(0, levels - 1)
        kind = rnd.choice(["texture", "noise", "two-tone"])
        rng = np.random.default_rng(done)
        if kind == "texture":
            img = _image(w, h, ch, done)
        elif kind == "noise":
            img = rng.integers(0, 256, (ch, h, w), dtype=np.uint8)
        else:
            img = np.full((ch, h, w), 40, dtype=np.uint8)
            img[:, h // 2:, w // 3:] = 200
        try:
            o = oraclebind.Oracle(w, h, levels, lc, trans=trans)
        except Exception:
            continue  # geometry below the codec's minimum
        want = refbind.compress(img, q, trans=trans, levels=levels, level_chg=lc)
        arenas = o.encode_image(img, q)
        got = capi.entropy_encode(w, h, ch, arenas.copy(), levels=levels, level_chg=lc, cap=4 * w * h * ch + 4096)
        assert got.tobytes() == bytes(want), (w, h, ch, q, levels, lc, trans, kind)
        keep = arenas.copy()
        hinted = capi.entropy_encode(w, h, ch, arenas, levels=levels, level_chg=lc, cap=4 * w * h * ch + 4096, hinted=True)
        assert hinted.tobytes() == bytes(want) and np.array_equal(arenas, keep), (w, h, ch, q, levels, lc, trans, kind)
        back = np.full(arenas.size, 0x11, dtype=np.uint8)
        capi.entropy_decode(w, h, ch, got, back, levels=levels, level_chg=lc)
        for p in range(ch):
            o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
        assert np.array_equal(back, arenas), (w, h, ch, q, levels, lc, trans, kind)
        done += 1


@pytest.mark.parametrize("fill", [0xFF, 0x55, 0xAA, 0x00, 0x5A])
@pytest.mark.parametrize("ch", [1, 3])
def test_pathological_payloads_are_rejected_not_crashing(fill, ch):
    """a payload without two consecutive zero bits used to run MuxReader::taboo() past its 32-entry
    tables (and corrupt streams could push the Golomb parameter past 31).  Such payloads must come back as an
    error (RIC_E_ARG) or decode to *something* -- never touch memory outside the tables."""
    w, h = 64, 48
    o = oraclebind.Oracle(w, h, 5)
    for size in (1, 7, 64, 4096):
        payload = np.full(size, fill, dtype=np.uint8)
        back = np.zeros(o.arena_bytes * ch, dtype=np.uint8)
        try:
            capi.entropy_decode(w, h, ch, payload, back)
        except capi.RicError as e:
            assert e.code == -1, e  # RIC_E_ARG
