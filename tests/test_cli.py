"""The command-line tool (rududu_image_codec_b200/ric_b200): PNM <-> .ric with the
reference tool's options (src/ric/ric.cpp:300-360)."""
import os
import subprocess

import numpy as np
import pytest

import oraclebind
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CLI = os.path.join(ROOT, "rududu_image_codec_b200", "ric_b200")


def _write_pnm(path, img):
    ch, h, w = img.shape
    with open(path, "wb") as f:
        f.write(b"P%d\n# synthetic\n%d %d\n255\n" % (6 if ch == 3 else 5, w, h))
        f.write(np.ascontiguousarray(img.transpose(1, 2, 0)).tobytes())


def _read_pnm(path):
    d = open(path, "rb").read()
    magic, dims, maxval, px = d.split(b"\n", 3)
    w, h = map(int, dims.split())
    ch = 3 if magic == b"P6" else 1
    assert maxval == b"255" and len(px) == w * h * ch
    return np.frombuffer(px, np.uint8).reshape(h, w, ch).transpose(2, 0, 1)


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)


def test_cli_is_built_and_has_no_cpu_path(tmp_path):
    assert os.access(CLI, os.X_OK), "run __graft_entry__.build()"
    assert _run("-h").returncode == 0
    assert _run().returncode == 1
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    p = tmp_path / "a.pgm"
    _write_pnm(p, synth_image(0, 64, 48, 1))
    r = _run("-i", str(p))
    assert r.returncode == 1 and "ric_create" in r.stderr  # fails loudly: the stage exists on the GPU only
    assert not (tmp_path / "a.ric").exists()


def _dither_ref(plane):
    """Independent restatement of ric.cpp:51-74 on a 12.4 fixed-point plane (int16 stores)."""
    h, w = plane.shape
    p = plane.astype(np.int64).reshape(-1).tolist()
    s16 = lambda v: ((v + 0x8000) & 0xFFFF) - 0x8000
    plain = lambda v: min(max(128 + ((v + 8) >> 4), 0), 255)
    o = 0
    for _ in range(h - 1):
        p[o] = plain(p[o])
        for i in range(1, w - 1):
            t = s16(p[o + i] + 8)
            q = t >> 4
            t = s16(t - (q << 4))
            p[o + i + 1] = s16(p[o + i + 1] + (t >> 1) - (t >> 4))
            p[o + i + w - 1] = s16(p[o + i + w - 1] + (t >> 3) + (t >> 4))
            p[o + i + w] = s16(p[o + i + w] + (t >> 2) + (t >> 4))
            p[o + i + w + 1] = s16(p[o + i + w + 1] + (t >> 4))
            p[o + i] = min(max(q + 128, 0), 255)
        o += w
        p[o - 1] = plain(p[o - 1])
    for i in range(w):
        p[o + i] = plain(p[o + i])
    return np.array(p, dtype=np.uint8).reshape(h, w)


@pytest.mark.gpu
@pytest.mark.parametrize("ch,q,trans", [(3, 9, None), (1, 5, None), (3, 0, None), (1, 13, 1)])
def test_cli_round_trip_matches_oracle(tmp_path, ch, q, trans):
    w, h = 250, 134
    img = synth_image(3, w, h, ch)
    src = tmp_path / ("in.ppm" if ch == 3 else "in.pgm")
    _write_pnm(src, img)
    args = ["-i", str(src), "-q", str(q)] + (["-t", str(trans)] if trans is not None else [])
    r = _run(*args)
    assert r.returncode == 0, r.stderr
    t = trans if trans is not None else (1 if q == 0 else 0)   # lossless defaults to 5/3 (ric.cpp:309)
    o = oraclebind.Oracle(w, h, 5, trans=t)
    arenas = o.encode_image(img, q)
    want = capi.header_write(w, h, q, int(ch == 3), t) + capi.entropy_encode(w, h, ch, arenas.copy()).tobytes()
    ric = tmp_path / "in.ric"                                   # default name: extension replaced
    assert ric.read_bytes() == want
    r = _run("-i", str(ric))
    assert r.returncode == 0, r.stderr
    for p in range(ch):
        o.unfold(arenas[p * o.arena_bytes:(p + 1) * o.arena_bytes])
    dec = _read_pnm(str(ric) + ".pnm")                           # default name: ".pnm" appended
    assert np.array_equal(dec, o.decode_image(arenas, ch, q))
    if q == 0:
        assert np.array_equal(dec, img)


@pytest.mark.gpu
def test_cli_dithered_gray_output(tmp_path):
    w, h, q = 96, 80, 13
    img = synth_image(5, w, h, 1)
    src = tmp_path / "g.pgm"
    _write_pnm(src, img)
    assert _run("-i", str(src), "-q", str(q), "-o", str(tmp_path / "g.ric")).returncode == 0
    r = _run("-i", str(tmp_path / "g.ric"), "-d", "-o", str(tmp_path / "d.pgm"))
    assert r.returncode == 0, r.stderr
    o = oraclebind.Oracle(w, h, 5)
    arena = o.encode_image(img, q)
    o.unfold(arena)
    Q, _ = oraclebind.plane_quant(q, 1, 0)
    o.tsuqi(arena, Q)
    plane = o.inverse(arena)
    assert np.array_equal(_read_pnm(tmp_path / "d.pgm")[0], _dither_ref(plane))
