"""GPU parity: the CUDA path (through the C ABI, via ctypes) against the oracle on the same inputs.

Bit-exact (integer path): every comparison is array_equal / CRC equality, tolerance 0.
The oracle (oracle/ric_oracle.c) is itself pinned against the compiled reference by
tests/test_oracle_vs_ref.py; tests/golden/kats.json holds reference-generated CRCs for the
BASELINE.json shapes.
"""
import numpy as np
import pytest

import oraclebind
import refbind
from refutil import crc, ref_payload_from_arenas
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image

pytestmark = pytest.mark.gpu

SHAPES = [(512, 512), (517, 389), (64, 48), (33, 47), (250, 131), (16, 16), (129, 130), (96, 17), (241, 64),
          (480, 270), (1000, 40)]
CFGS = [(5, 1, 0), (5, 0, 0), (6, 2, 0), (4, 3, 0), (5, 1, 1), (5, 0, 1)]


def _diff(o, a, b, nplanes=1):
    for p in range(nplanes):
        for i in range(o.nbands):
            f = o.info(i)
            x = o.band_view(a, i, p)[:, :f["dimx"]]
            y = o.band_view(b, i, p)[:, :f["dimx"]]
            if not np.array_equal(x, y):
                bad = np.argwhere(x != y)
                return "plane %d band %d (%dx%d): %d diffs, first at %s: got %d want %d" % (
                    p, i, f["dimx"], f["dimy"], len(bad), bad[0], x[tuple(bad[0])], y[tuple(bad[0])])
    return None


@pytest.mark.parametrize("w,h", SHAPES)
@pytest.mark.parametrize("levels,chg,trans", CFGS)
def test_transform_full_range(w, h, levels, chg, trans):
    rng = np.random.default_rng(w * 7919 + h + levels + 10 * trans)
    plane = rng.integers(-32768, 32768, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, levels, chg, trans=trans)
    with capi.Context(w, h, 1, levels, chg, trans=trans) as c:
        assert c.arena_bytes == o.arena_bytes
        for i in range(o.nbands):
            f, g = o.info(i), c.band(i)
            assert f == g
        got = c.transform(plane)
    assert _diff(o, got, o.forward(plane)) is None


@pytest.mark.parametrize("w,h", SHAPES)
@pytest.mark.parametrize("levels,chg,trans", [(5, 1, 0), (5, 0, 0), (6, 2, 0), (5, 1, 1)])
def test_transform_inv_full_range(w, h, levels, chg, trans):
    rng = np.random.default_rng(w * 31 + h + levels)
    o = oraclebind.Oracle(w, h, levels, chg, trans=trans)
    a = o.new_arena()
    for i in range(o.nbands):
        f = o.info(i)
        lo, hi = (-2 ** 20, 2 ** 20) if f["is_int"] else (-32768, 32768)
        o.band_view(a, i)[:, :f["dimx"]] = rng.integers(lo, hi, size=(f["dimy"], f["dimx"]))
    with capi.Context(w, h, 1, levels, chg, trans=trans) as c:
        got = c.transform_inv(a)
    assert np.array_equal(got, o.inverse(a, q1_quirk=1))


@pytest.mark.parametrize("Quant,lam", [(7, 3), (96, 36), (1000, 400), (6144, 3000), (30000, 12000)])
@pytest.mark.parametrize("w,h", [(203, 151), (512, 256)])
def test_quant_random_coefficients(Quant, lam, w, h):
    """Heavy-tailed planes: ties in the candidate ranking, every rank threshold, partial blocks,
    int32 coarse levels, and (large Quant) the int16 wrap paths."""
    rng = np.random.default_rng(Quant)
    plane = (rng.standard_cauchy(size=(h, w)) * Quant / 6).clip(-32768, 32767).astype(np.int16)
    o = oraclebind.Oracle(w, h, 5, 1)
    want = o.forward(plane)
    o.quant(want, Quant, lam)
    with capi.Context(w, h, 1, 5, 1) as c:
        c.transform(plane)
        got = c.quant(Quant, lam)
    assert _diff(o, got, want) is None


def test_tsuq_and_tsuqi():
    w, h = 250, 131
    rng = np.random.default_rng(5)
    plane = rng.integers(-3000, 3000, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, 5, 1)
    want = o.forward(plane)
    cnt = o.tsuq_all(want, 96, 0.7)
    with capi.Context(w, h, 1, 5, 1) as c:
        c.transform(plane)
        got, gcnt = c.tsuq(96, 0.7)
        assert gcnt == cnt
        assert _diff(o, got, want) is None
        o.tsuqi(want, 96)
        assert _diff(o, c.tsuqi(got, 96), want) is None


@pytest.mark.parametrize("w,h", [(512, 512), (517, 389), (250, 131), (33, 47), (129, 130), (1920, 1080)])
@pytest.mark.parametrize("q", [0, 1, 4, 9, 16, 31])
@pytest.mark.parametrize("ch", [1, 3])
def test_encode_stage(w, h, q, ch):
    img = synth_image(3, w, h, ch)
    trans = 1 if q == 0 else 0
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    want = o.encode_image(img, q)
    with capi.Context(w, h, ch, 5, trans=trans) as c:
        got = c.encode_u8(img[None], q)
    assert _diff(o, got, want, ch) is None
    assert np.array_equal(got, want)  # including the zeroed padding columns


@pytest.mark.parametrize("w,h", [(512, 512), (517, 389), (250, 131), (33, 47), (129, 130), (1920, 1080)])
@pytest.mark.parametrize("q", [0, 1, 9, 20, 31])
@pytest.mark.parametrize("ch", [1, 3])
def test_decode_stage(w, h, q, ch):
    img = synth_image(1, w, h, ch)
    trans = 1 if q == 0 else 0
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    arenas = o.encode_image(img, q)
    for p in range(ch):
        o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
    want = o.decode_image(arenas, ch, q)
    with capi.Context(w, h, ch, 5, trans=trans) as c:
        got = c.decode_u8(arenas, 1, q)[0]
    assert np.array_equal(got, want)
    if q == 0 and w % 2 == 0:
        assert np.array_equal(got, img)


def test_batch_matches_single():
    w, h, ch, n, q = 480, 272, 3, 5, 9
    imgs = np.stack([synth_image(i, w, h, ch) for i in range(n)])
    o = oraclebind.Oracle(w, h, 5)
    with capi.Context(w, h, ch, 5, max_batch=n) as c:
        got = c.encode_u8(imgs, q)
        for i in range(n):
            want = o.encode_image(imgs[i], q)
            assert np.array_equal(got[i * c.image_arena_bytes:(i + 1) * c.image_arena_bytes], want), i
        signed = got.copy()
        for i in range(n * ch):
            o.unfold(signed[i * o.arena_bytes:(i + 1) * o.arena_bytes])
        dec = c.decode_u8(signed, n, q)
        for i in range(n):
            want = o.decode_image(signed[i * c.image_arena_bytes:(i + 1) * c.image_arena_bytes], ch, q)
            assert np.array_equal(dec[i], want), i


def test_encode_after_decode_keeps_padding_zero():
    w, h, q = 250, 131, 9
    img = synth_image(0, w, h, 1)
    o = oraclebind.Oracle(w, h, 5)
    want = o.encode_image(img, q)
    with capi.Context(w, h, 1, 5) as c:
        junk = np.full(c.arena_bytes, 0x5A, dtype=np.uint8)
        c.decode_u8(junk, 1, q)
        assert np.array_equal(c.encode_u8(img[None], q), want)


def test_golden_full_size(golden):
    """BASELINE.json shapes at full size against reference-generated CRCs (tests/golden/kats.json)."""
    for k in golden["kats"]:
        w, h, ch, q = k["w"], k["h"], k["ch"], k["q"]
        img = synth_image(k["idx"], w, h, ch)
        assert crc(img) == k["src_crc"]
        o = oraclebind.Oracle(w, h, k["levels"], trans=k["trans"])
        with capi.Context(w, h, ch, k["levels"], trans=k["trans"]) as c:
            arenas = c.encode_u8(img[None], q)
            assert crc(arenas) == k["enc_arena_crc"], k
            for p in range(ch):
                o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
            assert crc(arenas) == k["dec_arena_crc"], k
            dec = c.decode_u8(arenas, 1, q)
            assert crc(dec) == k["dec8_crc"], k


def test_error_codes():
    L = capi.lib()
    with pytest.raises(capi.RicError) as e:
        capi.Context(8, 8)
    assert e.value.code == capi.E_ARG
    with pytest.raises(capi.RicError) as e:
        capi.Context(66, 64, trans=capi.HAAR)  # 33 columns at the second level
    assert e.value.code == capi.E_UNSUPPORTED
    with capi.Context(64, 64) as c:
        with pytest.raises(capi.RicError):
            c.encode_u8(np.zeros((2, 1, 64, 64), np.uint8), 9)  # n > max_batch
        with pytest.raises(capi.RicError):
            c.encode_u8(np.zeros((1, 1, 64, 64), np.uint8), 40)
    assert L.ric_last_error()


@pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libric_ref.so not built")
def test_ric_bitstream_bytes_from_gpu_bands(golden):
    """.ric payload parity: GPU-produced band arenas handed to the reference's own entropy coder
    (CBandCodec::pred/tree + CMuxCodec, unmodified) give the reference's bitstream byte for byte
    (SURVEY.md Appendix C payload size + CRC32), for every golden shape that is cheap to code on the CPU."""
    for k in golden["kats"]:
        w, h, ch, q = k["w"], k["h"], k["ch"], k["q"]
        if w * h * ch > 3840 * 2160 * 3 or k["trans"] != 0 or (w * h * ch > 1 << 22 and q not in (9, 27)):
            continue
        img = synth_image(k["idx"], w, h, ch)
        o = oraclebind.Oracle(w, h, k["levels"], trans=k["trans"])
        with capi.Context(w, h, ch, k["levels"], trans=k["trans"]) as c:
            arenas = c.encode_u8(img[None], q)
        payload = ref_payload_from_arenas(o, arenas, ch, k["trans"])
        assert (len(payload), crc(payload)) == (k["payload_bytes"], k["payload_crc"]), k


@pytest.mark.parametrize("w,h", [(239, 40), (240, 40), (247, 33), (248, 33), (249, 64), (479, 24), (481, 24), (720, 135),
                                  (2000, 16), (16, 2000), (65535, 16), (16, 65535)])
def test_strip_and_segment_boundaries(w, h):
    """Widths straddling the 240-column strip / 8-column lane grid, heights straddling segment and
    4x4-block-row boundaries, extreme aspect ratios up to the u16 header limit (ric.cpp:150-153)."""
    img = synth_image(7, w, h, 1)
    o = oraclebind.Oracle(w, h, 5)
    want = o.encode_image(img, 9)
    with capi.Context(w, h, 1, 5) as c:
        got = c.encode_u8(img[None], 9)
        assert _diff(o, got, want) is None
        o.unfold(want)
        assert np.array_equal(c.decode_u8(want, 1, 9)[0], o.decode_image(want, 1, 9))


@pytest.mark.parametrize("align", [32, 64, 128])
def test_alignments(align):
    w, h = 250, 130
    rng = np.random.default_rng(align)
    plane = rng.integers(-32768, 32768, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, 5, 1, align=align)
    with capi.Context(w, h, 1, 5, 1, align=align) as c:
        got = c.transform(plane)
        assert _diff(o, got, o.forward(plane)) is None
        assert np.array_equal(c.transform_inv(got), plane)  # lossless identity of the 9/7 lifting


def test_device_api_with_padded_pitch():
    """ric_*_device with a caller pitch larger than the width (rows padded to 512 bytes)."""
    import torch
    w, h, ch, q, pitch = 333, 77, 3, 9, 512
    img = synth_image(2, w, h, ch)
    o = oraclebind.Oracle(w, h, 5)
    want = o.encode_image(img, q)
    with capi.Context(w, h, ch, 5) as c:
        src = torch.zeros((ch, h, pitch), dtype=torch.uint8, device="cuda")
        src[:, :, :w] = torch.from_numpy(img).cuda()
        ar = torch.zeros(c.image_arena_bytes, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        c.encode_u8_device(src.data_ptr(), pitch, 1, q, ar.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(ar.cpu().numpy(), want)
        for p in range(ch):
            o.unfold(want[p * o.arena_bytes:(p + 1) * o.arena_bytes])
        dst = torch.zeros((ch, h, pitch), dtype=torch.uint8, device="cuda")
        c.decode_u8_device(torch.from_numpy(want).cuda().data_ptr(), 1, q, dst.data_ptr(), pitch, st)
        torch.cuda.synchronize()
        assert np.array_equal(dst[:, :, :w].cpu().numpy(), o.decode_image(want, ch, q))


def test_stream_api_chunk_callbacks():
    """ric_encode_u8_stream / ric_decode_u8_stream: every chunk's data is complete when its callback fires."""
    w, h, ch, n, q = 320, 200, 3, 9, 9
    imgs = np.stack([synth_image(i, w, h, ch) for i in range(n)])
    o = oraclebind.Oracle(w, h, 5)
    want = [o.encode_image(imgs[i], q) for i in range(n)]
    with capi.Context(w, h, ch, 5, max_batch=n) as c:
        out = np.zeros(n * c.image_arena_bytes, dtype=np.uint8)
        seen, ok = [], []

        def done(first, cnt):
            seen.append((first, cnt))
            for i in range(first, first + cnt):
                ok.append(np.array_equal(out[i * c.image_arena_bytes:(i + 1) * c.image_arena_bytes], want[i]))

        c.encode_u8_stream(imgs, q, out, done)
        c.sync()
        assert sorted(i for f, k in seen for i in range(f, f + k)) == list(range(n))
        assert len(seen) > 1 and all(ok) and len(ok) == n
        signed = out.copy()
        for i in range(n * ch):
            o.unfold(signed[i * o.arena_bytes:(i + 1) * o.arena_bytes])
        dec = np.zeros((n, ch, h, w), dtype=np.uint8)
        seen2 = []
        c.decode_u8_stream(signed, n, q, dec, lambda f, k: seen2.append((f, k)))
        c.sync()
        assert sum(k for _, k in seen2) == n
        for i in range(n):
            assert np.array_equal(dec[i], o.decode_image(signed[i * c.image_arena_bytes:(i + 1) * c.image_arena_bytes], ch, q))


def test_contexts_are_independent_across_threads():
    """One ric_ctx per thread (the reference's threading rule): concurrent encodes on
    different contexts do not interfere."""
    import threading
    shapes = [(320, 200, 3), (250, 131, 1), (512, 256, 3), (129, 130, 1)]
    results = {}

    def work(k):
        w, h, ch = shapes[k]
        img = synth_image(k, w, h, ch)
        with capi.Context(w, h, ch, 5) as c:
            for _ in range(5):
                results[k] = c.encode_u8(img[None], 9)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(shapes))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k, (w, h, ch) in enumerate(shapes):
        want = oraclebind.Oracle(w, h, 5).encode_image(synth_image(k, w, h, ch), 9)
        assert np.array_equal(results[k], want), k


@pytest.mark.parametrize("w,h,levels,chg", [(512, 512, 5, 1), (320, 192, 5, 1), (64, 96, 3, 0), (640, 352, 5, 0), (1920, 1088, 5, 1)])
def test_haar(w, h, levels, chg):
    rng = np.random.default_rng(w + h)
    plane = rng.integers(-32768, 32768, size=(h, w), dtype=np.int16)
    o = oraclebind.Oracle(w, h, levels, chg, trans=2)
    with capi.Context(w, h, 1, levels, chg, trans=capi.HAAR) as c:
        got = c.transform(plane)
        assert _diff(o, got, o.forward(plane)) is None
        assert np.array_equal(c.transform_inv(got), plane)
    # whole stage, lossless and lossy, RGB
    img = synth_image(4, w, h, 3)
    for q in (0, 9):
        want = o.encode_image(img, q)
        with capi.Context(w, h, 3, levels, chg, trans=capi.HAAR) as c:
            got = c.encode_u8(img[None], q)
            assert np.array_equal(got, want)
            for p in range(3):
                o.unfold(want[p * o.arena_bytes:(p + 1) * o.arena_bytes])
            dec = c.decode_u8(want, 1, q)[0]
        assert np.array_equal(dec, o.decode_image(want, 3, q))
        if q == 0:
            assert np.array_equal(dec, img)


def test_haar_odd_sizes_are_refused():
    with pytest.raises(capi.RicError) as e:
        capi.Context(1920, 1080, trans=capi.HAAR)  # 1080/8 = 135 is odd at level 4
    assert e.value.code == capi.E_UNSUPPORTED


def _adversarial_images(w, h, ch, rng):
    yy, xx = np.mgrid[0:h, 0:w]
    chk = (((xx + yy) & 1) * 255).astype(np.uint8)
    yield "uniform noise", rng.integers(0, 256, size=(ch, h, w), dtype=np.uint8)
    yield "checkerboard", np.stack([chk if c != 1 else 255 - chk for c in range(ch)])
    yield "column stripes", np.stack([((xx & 1) * 255).astype(np.uint8)] * ch)
    yield "row stripes 2", np.stack([(((yy >> 1) & 1) * 255).astype(np.uint8) if c == 0 else rng.integers(0, 2, size=(h, w), dtype=np.uint8) * 255
                                     for c in range(ch)])
    yield "saturated", np.full((ch, h, w), 255, dtype=np.uint8)


@pytest.mark.parametrize("ch", [1, 3])
@pytest.mark.parametrize("q", [0, 1, 9, 31])
def test_adversarial_pixels(ch, q):
    """Worst-case 8-bit inputs (maximum-amplitude alternation): the deepest levels really wrap around
    in int16 here, and the level-0 row pass runs without truncation (proved bound) -- both must match
    the reference arithmetic bit for bit."""
    w, h = 496, 264
    trans = 1 if q == 0 else 0
    rng = np.random.default_rng(q * 10 + ch)
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    with capi.Context(w, h, ch, 5, trans=trans) as c:
        for name, img in _adversarial_images(w, h, ch, rng):
            img = np.ascontiguousarray(img)
            want = o.encode_image(img, q)
            got = c.encode_u8(img[None], q)
            assert _diff(o, got, want, ch) is None, name
            for p in range(ch):
                o.unfold(want[p * o.arena_bytes:(p + 1) * o.arena_bytes])
            assert np.array_equal(c.decode_u8(want, 1, q)[0], o.decode_image(want, ch, q)), name


@pytest.mark.parametrize("ch", [1, 3])
@pytest.mark.parametrize("q,trans", [(9, 0), (31, 0), (0, 1), (9, 2), (0, 0)])
def test_decode_stage_full_range_arenas(ch, q, trans):
    """Decode stage on arbitrary full-range coefficient arenas (what a corrupt or hostile bitstream
    could produce): dequantisation products, lifting and the colour transform all wrap as the
    reference's short arithmetic does."""
    w, h = 320, 192
    rng = np.random.default_rng(100 * q + trans + ch)
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    arenas = np.zeros(ch * o.arena_bytes, dtype=np.uint8)
    for p in range(ch):
        for i in range(o.nbands):
            f = o.info(i)
            lo, hi = (-2 ** 31, 2 ** 31) if f["is_int"] else (-32768, 32768)
            o.band_view(arenas, i, p)[:, :f["dimx"]] = rng.integers(lo, hi, size=(f["dimy"], f["dimx"]))
    with capi.Context(w, h, ch, 5, trans=trans) as c:
        got = c.decode_u8(arenas, 1, q)[0]
    assert np.array_equal(got, o.decode_image(arenas, ch, q))


def test_bench_decode_input_preparation_matches_oracle_unfold():
    """bench.py prepares its decode input with torch ops (it may not call the oracle): check that
    helper against rico_unfold, i.e. against what the reference's DecodeBand leaves in the bands."""
    import importlib.util
    import os
    import torch
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w, h, ch, q = 320, 200, 3, 9
    img = synth_image(0, w, h, ch)
    o = oraclebind.Oracle(w, h, 5)
    with capi.Context(w, h, ch, 5) as c:
        arenas = c.encode_u8(img[None], q)
        t = torch.from_numpy(arenas.copy()).cuda()
        bench.unfold_arenas_(c, t, 1)
        got = t.cpu().numpy()
    for p in range(ch):
        o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
    assert np.array_equal(got, arenas)


def test_ric_files_match_golden(golden):
    """Whole .ric files through ric_compress_u8 (GPU stage + product entropy stage on host threads): header
    bytes, payload size and CRC32 equal the reference's known answers; ric_decompress_u8
    of those files gives the reference decoder's pixels.  Every known answer, the 8192 x 8192 gray image
    (17 749 195 bytes, CRC 6ca7ca96) and the whole 4K quantiser sweep included."""
    assert any(k["w"] == 8192 for k in golden["kats"]) and sum(k["w"] == 3840 for k in golden["kats"]) >= 8
    for k in golden["kats"]:
        w, h, ch, q = k["w"], k["h"], k["ch"], k["q"]
        img = synth_image(k["idx"], w, h, ch)
        with capi.Context(w, h, ch, k["levels"], trans=k["trans"]) as c:
            f = c.compress_u8(img[None], q)[0]
            assert f[:9] == capi.header_write(w, h, q, int(ch == 3), k["trans"])
            assert (len(f) - 9, crc(np.frombuffer(f[9:], np.uint8))) == (k["payload_bytes"], k["payload_crc"]), k
            dec = c.decompress_u8([f])
            assert crc(dec) == k["dec8_crc"], k


@pytest.mark.parametrize("ch,q,threads", [(3, 9, 0), (1, 5, 3), (3, 0, 2)])
def test_ric_file_batch(ch, q, threads):
    """A batch of different images: every file equals oracle arenas + entropy stage coded one by one, files come
    back in order whatever thread finished first, and decompress inverts to the oracle's decoder output."""
    w, h, n = 400, 300, 13
    trans = 1 if q == 0 else 0
    imgs = np.stack([synth_image(20 + i, w, h, ch) for i in range(n)])
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    with capi.Context(w, h, ch, 5, trans=trans, max_batch=n) as c:
        files = c.compress_u8(imgs, q, threads=threads)
        for i in range(n):
            arenas = o.encode_image(imgs[i], q)
            want = capi.entropy_encode(w, h, ch, arenas.copy())
            assert files[i][:9] == capi.header_write(w, h, q, int(ch == 3), trans)
            assert files[i][9:] == want.tobytes(), i
        dec = c.decompress_u8(files, threads=threads)
        for i in range(n):
            arenas = o.encode_image(imgs[i], q)
            for p in range(ch):
                o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
            assert np.array_equal(dec[i], o.decode_image(arenas, ch, q)), i
        # a second call on the same context reuses the staging arenas
        assert c.compress_u8(imgs[:5], q, threads=1) == files[:5]
        # the same files with the entropy stage on the device, and back
        assert c.compress_u8(imgs, q, entropy_on_device=True) == files
        assert np.array_equal(c.decompress_u8(files, entropy_on_device=True), dec)


def test_ric_file_errors():
    w, h = 128, 96
    img = synth_image(1, w, h, 3)
    with capi.Context(w, h, 3, 5, max_batch=2) as c:
        with pytest.raises(capi.RicError) as e:
            c.compress_u8(img[None], 1, stride=64)
        assert e.value.code == capi.E_NOMEM
        f9 = c.compress_u8(img[None], 9)[0]
        f5 = c.compress_u8(img[None], 5)[0]
        with pytest.raises(capi.RicError) as e:
            c.decompress_u8([f9, f5])          # different quantiser index in one batch
        assert e.value.code == capi.E_ARG
        with pytest.raises(capi.RicError):
            c.decompress_u8([b"RUD1" + f9[4:]])  # bad magic
        with pytest.raises(capi.RicError):
            c.decompress_u8([f9[:len(f9) // 3]])  # truncated payload
        with pytest.raises(capi.RicError) as e:
            c.compress_u8(img[None], 1, stride=64, entropy_on_device=True)
        assert e.value.code == capi.E_NOMEM
        with pytest.raises(capi.RicError):
            c.decompress_u8([f9[:len(f9) // 3]], entropy_on_device=True)
        assert np.array_equal(c.decompress_u8([f9], entropy_on_device=True), c.decompress_u8([f9]))
    with capi.Context(w, h, 1, 5) as c:
        with pytest.raises(capi.RicError) as e:
            c.decompress_u8([f9])                # colour file, gray context
        assert e.value.code == capi.E_ARG


@pytest.mark.parametrize("ch,q,trans,w,h", [(3, 9, 0, 400, 300), (1, 5, 0, 250, 134), (3, 0, 1, 128, 96)])
def test_entropy_stage_on_device(ch, q, trans, w, h):
    """The device entropy stage (one image per warp, same coder source) against the host one: payloads byte
    for byte from device-resident encode-stage arenas, and back to the signed arenas the decode stage reads."""
    import torch
    n = 5
    imgs = np.stack([synth_image(40 + i, w, h, ch) for i in range(n)])
    o = oraclebind.Oracle(w, h, 5, trans=trans)
    stride = w * h * ch * 2 + 4096
    with capi.Context(w, h, ch, 5, trans=trans, max_batch=n) as c:
        st = torch.cuda.current_stream().cuda_stream
        pitch = (w + 7) & ~7
        src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
        src[..., :w] = torch.from_numpy(imgs).cuda()
        ar = torch.zeros(n * c.image_arena_bytes, dtype=torch.uint8, device="cuda")
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
        out = torch.zeros(n * stride, dtype=torch.uint8, device="cuda")
        sizes = torch.zeros(n, dtype=torch.int64, device="cuda")
        c.entropy_encode_device(ar.data_ptr(), n, out.data_ptr(), stride, sizes.data_ptr(), st)
        torch.cuda.synchronize()
        out_h, sizes_h = out.cpu().numpy().reshape(n, stride), sizes.cpu().numpy()
        signed = []
        for i in range(n):
            arenas = o.encode_image(imgs[i], q)
            want = capi.entropy_encode(w, h, ch, arenas.copy())
            assert sizes_h[i] == want.size and out_h[i, :want.size].tobytes() == want.tobytes(), i
            for p in range(ch):
                o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
            signed.append(arenas)
        back = torch.full((n * c.image_arena_bytes,), 0x5A, dtype=torch.uint8, device="cuda")
        status = torch.full((n,), 7, dtype=torch.int32, device="cuda")
        c.entropy_decode_device(out.data_ptr(), stride, sizes.data_ptr(), n, back.data_ptr(), st, d_status=status.data_ptr())
        assert (status.cpu().numpy() == 0).all()
        dst = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
        c.decode_u8_device(back.data_ptr(), n, q, dst.data_ptr(), pitch, st)
        torch.cuda.synchronize()
        assert np.array_equal(back.cpu().numpy(), np.concatenate(signed))
        for i in range(n):
            assert np.array_equal(dst[i, :, :, :w].cpu().numpy(), o.decode_image(signed[i], ch, q)), i
        # too small a slot is reported per image, not written past
        small = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
        c.entropy_encode_device(ar.data_ptr(), n, small.data_ptr(), 64, sizes.data_ptr(), st)
        torch.cuda.synchronize()
        assert (sizes.cpu().numpy() == -1).all()
        # ... and the decoder reports bad inputs per image: a negative length, a truncated payload, a
        # payload of 0xFF bytes (no code of the format has that shape); the good images still decode
        if n >= 3:
            bad_sizes = torch.from_numpy(sizes_h.copy()).cuda()
            bad_sizes[0] = -1
            bad_sizes[1] = max(int(sizes_h[1]) // 3, 1)
            junk = out.clone()
            junk.view(n, stride)[2, :] = 0xFF
            c.entropy_decode_device(junk.data_ptr(), stride, bad_sizes.data_ptr(), n, back.data_ptr(), st, d_status=status.data_ptr())
            torch.cuda.synchronize()
            stat = status.cpu().numpy()
            assert stat[0] == 1 and stat[1] == 1 and stat[2] == 1 and (stat[3:] == 0).all(), stat
            per = c.image_arena_bytes
            assert np.array_equal(back[3 * per:].cpu().numpy(), np.concatenate(signed[3:])) if n > 3 else True


def test_ric_file_many_workers():
    """More host workers than images left to hand out, repeatedly: every worker must come home
    (regression: a worker waiting for an image that another one then took used to wait forever)."""
    w, h, n = 64, 48, 40
    imgs = np.stack([synth_image(i, w, h, 1) for i in range(n)])
    with capi.Context(w, h, 1, 5, max_batch=n) as c:
        first = c.compress_u8(imgs, 9, threads=16)
        for _ in range(15):
            assert c.compress_u8(imgs, 9, threads=16) == first
        assert np.array_equal(c.decompress_u8(first, threads=16), c.decompress_u8(first, entropy_on_device=True))


def test_entropy_stage_on_device_plain_walker(monkeypatch):
    """RIC_ENTROPY_PLAIN=1 selects the plain device walker (no pre-pass; consumes the arenas like the host
    stage): same payloads as the default hinted encoder."""
    import torch
    w, h, ch, q, n = 320, 200, 3, 9, 3
    imgs = np.stack([synth_image(60 + i, w, h, ch) for i in range(n)])
    stride = w * h * ch * 2
    with capi.Context(w, h, ch, 5, max_batch=n) as c:
        st = torch.cuda.current_stream().cuda_stream
        pitch = (w + 7) & ~7
        src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
        src[..., :w] = torch.from_numpy(imgs).cuda()
        ar = torch.zeros(n * c.image_arena_bytes, dtype=torch.uint8, device="cuda")
        outs = []
        for plain in (False, True):
            if plain:
                monkeypatch.setenv("RIC_ENTROPY_PLAIN", "1")
            c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
            before = ar.clone()
            out = torch.zeros(n * stride, dtype=torch.uint8, device="cuda")
            sizes = torch.zeros(n, dtype=torch.int64, device="cuda")
            c.entropy_encode_device(ar.data_ptr(), n, out.data_ptr(), stride, sizes.data_ptr(), st)
            torch.cuda.synchronize()
            assert torch.equal(ar, before) != plain  # the hinted coder leaves the arenas alone, the plain one consumes them
            outs.append((out.cpu().numpy(), sizes.cpu().numpy()))
        assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0])
