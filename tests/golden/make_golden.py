"""Regenerates tests/golden/kats.json from the compiled reference (oracle/_ref/libric_ref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle`):
    python tests/golden/make_golden.py
Each entry: the SURVEY.md Appendix C known answers (src/payload/dec8 CRC32) plus the CRC32 of the
quantised band arenas (canonical layout, padding columns zero) the reference's quantiser half of
CodeBand produces -- the exact bytes the CUDA encode stage must emit -- and of the signed arenas
(DecodeBand output) the decode stage consumes.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import oraclebind  # noqa: E402
import refbind  # noqa: E402
from refutil import crc, ref_encode_arenas  # noqa: E402
from rududu_image_codec_b200.synth import synth_image  # noqa: E402

CASES = [  # w, h, ch, q, trans, levels, idx
    (512, 512, 1, 9, 0, 5, 0),
    (3840, 2160, 3, 9, 0, 5, 0),
    (8192, 8192, 1, 9, 0, 6, 0),
    (1920, 1080, 3, 9, 0, 5, 0),
    (1920, 1080, 3, 9, 0, 5, 1),
    (3840, 2160, 3, 1, 0, 5, 0),
    (3840, 2160, 3, 5, 0, 5, 0),
    (3840, 2160, 3, 13, 0, 5, 0),
    (3840, 2160, 3, 20, 0, 5, 0),
    (3840, 2160, 3, 27, 0, 5, 0),
    (3840, 2160, 3, 31, 0, 5, 0),
    (3840, 2160, 3, 0, 1, 5, 0),
    (512, 512, 1, 0, 1, 5, 0),
    (517, 389, 3, 9, 0, 5, 2),
    (246, 131, 1, 4, 0, 5, 3),
]

out = []
for (w, h, ch, q, trans, levels, idx) in CASES:
    img = synth_image(idx, w, h, ch)
    p = refbind.compress(img, q, trans, levels)
    d = refbind.decompress(p, w, h, ch, q, trans, levels)
    o, arenas = ref_encode_arenas(img, q, levels, trans=trans)
    signed = arenas.copy()
    for c in range(ch):
        o.unfold(signed[c * o.arena_bytes:(c + 1) * o.arena_bytes])
    e = dict(w=w, h=h, ch=ch, q=q, trans=trans, levels=levels, idx=idx, src_crc=crc(img),
             payload_bytes=int(len(p)), payload_crc=crc(p), dec8_crc=crc(d),
             arena_bytes=int(o.arena_bytes), enc_arena_crc=crc(arenas), dec_arena_crc=crc(signed))
    print(e, flush=True)
    out.append(e)
with open(os.path.join(HERE, "kats.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_golden.py", "kats": out}, f, indent=1)
