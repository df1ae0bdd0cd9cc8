"""Time the device entropy stage on a large batch (1080p RGB, BASELINE configs[3] shape) next to the host one.
usage: python scripts/gpu_entropy_time.py [n_images] [width height]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rududu_image_codec_b200 import capi  # noqa: E402
from rududu_image_codec_b200.synth import synth_image  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
ch, q, distinct = 3, 9, 8
imgs = np.stack([synth_image(i, w, h, ch) for i in range(distinct)])
stride = w * h * ch // 4
with capi.Context(w, h, ch, 5, max_batch=n) as c:
    st = torch.cuda.current_stream().cuda_stream
    pitch = (w + 7) & ~7
    src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
    src[..., :w] = torch.from_numpy(imgs).cuda()[torch.arange(n) % distinct]
    ar = torch.zeros(n * c.image_arena_bytes, dtype=torch.uint8, device="cuda")
    out = torch.zeros(n * stride, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(n, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(2):
        ev[0].record()
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
        ev[1].record()
        c.entropy_encode_device(ar.data_ptr(), n, out.data_ptr(), stride, sizes.data_ptr(), st)
        ev[2].record()
        c.entropy_decode_device(out.data_ptr(), stride, sizes.data_ptr(), n, ar.data_ptr(), st)
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        print("n=%d %dx%d: encode stage %.2f ms, entropy encode %.1f ms (%.1f Mpixel/s), entropy decode %.1f ms (%.1f Mpixel/s), "
              "mean payload %.0f B" % (n, w, h, t[0], t[1], n * w * h / t[1] / 1e3, t[2], n * w * h / t[2] / 1e3,
                                       float(sizes.float().mean())), flush=True)
    # host stage on the same arenas, one image, for the per-thread rate
    c.encode_u8_device(src.data_ptr(), pitch, 1, q, ar.data_ptr(), st)
    torch.cuda.synchronize()
    a = ar[:c.image_arena_bytes].cpu().numpy()
    t0 = time.perf_counter()
    p = capi.entropy_encode(w, h, ch, a)
    t1 = time.perf_counter()
    assert p.size == int(sizes[0]) and p.tobytes() == out[:p.size].cpu().numpy().tobytes()
    print("host entropy encode, one thread: %.1f ms per image (%.1f Mpixel/s per thread); device payload 0 identical"
          % (1e3 * (t1 - t0), w * h / (t1 - t0) / 1e6))
