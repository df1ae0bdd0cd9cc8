"""Small encode+decode workload for compute-sanitizer (odd sizes, edges, int levels, RGB + gray)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image
for (w, h, ch, lv, n) in [(517, 389, 3, 5, 2), (250, 131, 1, 6, 1), (1000, 40, 3, 5, 3)]:
    imgs = np.stack([synth_image(i, w, h, ch) for i in range(n)])
    with capi.Context(w, h, ch, lv, max_batch=n) as c:
        a = c.encode_u8(imgs, 9)
        d = c.decode_u8(a, n, 9)
        print(w, h, ch, lv, n, int(a.sum()), int(d.sum()))
print("done")
