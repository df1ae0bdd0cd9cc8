"""The C++ class shim (include/rududu_b200/wavelet2d.h), driven the way ric.cpp drives the reference."""
import os
import subprocess

import numpy as np
import pytest

import oraclebind
from rududu_image_codec_b200.synth import synth_image

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "rududu_image_codec_b200")


def _build(tmp_path, name="shim_test"):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-L" + PKG, "-lrududu_b200",
                           "-Wl,-rpath," + PKG])
    return exe


def test_codec_call_sequence_compiles(tmp_path):
    """ric.cpp's CompressImage / DecompressImage body compiles unchanged against the shim classes."""
    _build(tmp_path, "codec_test")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,ch,q", [(384, 256, 3, 9), (250, 134, 1, 5)])
def test_codec_call_sequence_matches_oracle(tmp_path, w, h, ch, q):
    """CMuxCodec + Transform/CodeBand per plane + endCoding, then DecodeBand/TSUQi/TransformI, through the
    shim: payload = oracle arenas through the entropy stage, planes = oracle inverse."""
    from rududu_image_codec_b200 import capi
    exe = _build(tmp_path, "codec_test")
    img = synth_image(2, w, h, ch)
    planes = oraclebind.colour_fwd(img, q)
    planes.tofile(tmp_path / "planes.s16")
    r = subprocess.run([exe, str(w), str(h), str(ch), str(q), str(tmp_path / "planes.s16"), str(tmp_path / "payload.bin"),
                        str(tmp_path / "out.s16")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o = oraclebind.Oracle(w, h, 5)
    arenas = o.encode_image(img, q)
    want = capi.entropy_encode(w, h, ch, arenas.copy())
    assert np.fromfile(tmp_path / "payload.bin", dtype=np.uint8).tobytes() == want.tobytes()
    got = np.fromfile(tmp_path / "out.s16", dtype=np.int16).reshape(ch, h, w)
    for p in range(ch):
        a = arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes]
        o.unfold(a)
        if q:
            o.tsuqi(a, oraclebind.plane_quant(q, ch, p)[0])
        assert np.array_equal(got[p], o.inverse(a)), p


def test_landed_queue_out_of_order_chunks(tmp_path):
    """ric_compress_u8's hand-off to its entropy workers: chunks land out of order (three streams); no image may be
    handed out before its own chunk has landed (regression: the queue once summarised arrivals as a high-water mark)."""
    exe = str(tmp_path / "landed_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "landed_test.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = subprocess.run([exe, "64", "64", "5", "96", "36", "/dev/null", "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(512, 512), (517, 389)])
def test_shim_matches_oracle(tmp_path, w, h):
    exe = _build(tmp_path)
    q = 9
    plane = oraclebind.colour_fwd(synth_image(0, w, h, 1), q)[0]
    Quant, lam = oraclebind.plane_quant(q, 1, 0)
    o = oraclebind.Oracle(w, h, 5)
    want = o.forward(plane)
    o.quant(want, Quant, lam)
    signed = want.copy()
    o.unfold(signed)
    plane.tofile(tmp_path / "plane.s16")
    signed.tofile(tmp_path / "signed.bin")
    r = subprocess.run([exe, str(w), str(h), "5", str(Quant), str(lam), str(tmp_path / "plane.s16"),
                        str(tmp_path / "arena.bin"), str(tmp_path / "signed.bin"), str(tmp_path / "out.s16")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "arena.bin", dtype=np.uint8), want)
    o.tsuqi(signed, Quant)
    assert np.array_equal(np.fromfile(tmp_path / "out.s16", dtype=np.int16).reshape(h, w), o.inverse(signed))


@pytest.mark.gpu
def test_streaming_pipeline_with_reference_entropy_coder(tmp_path):
    """configs[3] in miniature: GPU encode of a batch overlapped with host threads running the
    reference's own entropy coder on the chunks as they land; every .ric payload must equal the
    reference's monolithic CompressImage output."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "libric_ref.so")):
        pytest.skip("oracle/_ref not built")
    exe = str(tmp_path / "pipeline_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "pipeline_test.cpp"), "-L" + PKG, "-lrududu_b200",
                           "-L" + ref_dir, "-lric_ref", "-Wl,-rpath," + PKG, "-Wl,-rpath," + ref_dir])
    w, h, n, q = 480, 272, 13, 9
    imgs = np.stack([synth_image(i, w, h, 3) for i in range(n)])
    imgs.tofile(tmp_path / "imgs.u8")
    r = subprocess.run([exe, str(w), str(h), str(n), str(q), str(tmp_path / "imgs.u8")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,levels", [(512, 384, 5), (517, 389, 5), (130, 70, 2)])
def test_per_band_methods_match_oracle(tmp_path, w, h, levels):
    """CBandCodec::buildTree<high_band, C>, CBand::TSUQ<C> / TSUQi<C> / Init and SetWeight(t, baseWeight) of the shim
    (src/lib/bandcodec.h:42, src/lib/band.h:61-107, src/lib/wavelet2d.h:36), called band by band the way
    CWavelet2D::CodeBand / TSUQ / TSUQi call them, against the oracle's whole-plane functions."""
    exe = _build(tmp_path, "band_api_test")
    q = 9
    plane = oraclebind.colour_fwd(synth_image(4, w, h, 1), q)[0]
    Quant, lam = oraclebind.plane_quant(q, 1, 0)
    o = oraclebind.Oracle(w, h, levels)
    raw = o.forward(plane)
    want_bt = raw.copy()
    o.quant(want_bt, Quant, lam)
    signed = want_bt.copy()
    o.unfold(signed)
    plane.tofile(tmp_path / "plane.s16")
    signed.tofile(tmp_path / "signed.bin")
    pre = str(tmp_path / "o_")
    r = subprocess.run([exe, str(w), str(h), str(levels), str(Quant), str(lam), str(tmp_path / "plane.s16"),
                        str(tmp_path / "signed.bin"), pre], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(l.split(None, 1) for l in r.stdout.strip().splitlines() if " " in l)
    rd = lambda name: np.fromfile(pre + name + ".bin", dtype=np.uint8)
    assert np.array_equal(rd("bt"), want_bt)
    ll = o.band_view(want_bt, o.nbands - 1)[:, :o.info(o.nbands - 1)["dimx"]]
    f = out["ll_count"].split()
    assert (int(f[2]), int(f[4])) == (min(int(ll.min()), 0), max(int(ll.max()), 0))
    want_t = raw.copy()
    n = o.tsuq_all(want_t, Quant, 0.7)
    assert np.array_equal(rd("tsuq"), want_t)
    assert int(out["tsuq_count"].split()[0]) == n
    want_i = signed.copy()
    o.tsuqi(want_i, Quant)
    assert np.array_equal(rd("tsuqi"), want_i)
    want_bw = raw.copy()
    o.quant(want_bw, 2 * Quant, 2 * lam)  # every weight halved == Quant and lambda doubled (powers of two are exact in float)
    assert np.array_equal(rd("bw"), want_bw)
