"""Quick device-resident timing of the encode/decode stages (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image

def run(w, h, ch, levels, n, q=9, iters=20):
    img = synth_image(0, w, h, ch)
    c = capi.Context(w, h, ch, levels, max_batch=n)
    pitch = (w + 15) & ~7
    src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :, :w] = torch.from_numpy(img).cuda()[None]
    ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
    dst = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    te = e0.elapsed_time(e1) / iters
    for _ in range(3):
        c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
    e1.record(); torch.cuda.synchronize()
    td = e0.elapsed_time(e1) / iters
    S = w * h * ch * n
    print("%dx%dx%d n=%d L%d q=%d: encode %.3f ms (%.1f Gsample/s, %.0f GB/s alg) decode %.3f ms (%.1f Gsample/s, %.0f GB/s alg)" % (
        w, h, ch, n, levels, q, te, S / te / 1e6, 3 * S / te / 1e6, td, S / td / 1e6, 3 * S / td / 1e6), flush=True)
    c.close()

if __name__ == "__main__":
    run(3840, 2160, 3, 5, 1)
    run(8192, 8192, 1, 6, 1)
    run(1920, 1080, 3, 5, 64)
    run(1920, 1080, 3, 5, 64, q=20)
