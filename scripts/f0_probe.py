"""Development probe: level times and packed-path statistics of the encode stage (run with RIC_FWD0=0/1)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image

def run(w, h, ch, levels, n, q=9):
    imgs = np.stack([synth_image(i, w, h, ch) for i in range(min(n, 4))])
    c = capi.Context(w, h, ch, levels, max_batch=n)
    pitch = (w + 15) & ~15
    src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :, :w] = torch.from_numpy(imgs).cuda().repeat((n + 3) // 4, 1, 1, 1)[:n]
    ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    c.set_profiling(True)
    for _ in range(3):
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    torch.cuda.synchronize()
    c.path_stats()
    c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    torch.cuda.synchronize()
    print("%dx%dx%d n=%d q=%d FWD0=%s: level ms %s path stats %s" % (w, h, ch, n, q, os.environ.get("RIC_FWD0", "1"),
          ["%.3f" % t for t in c.level_times(0)], c.path_stats()), flush=True)
    c.close()

if __name__ == "__main__":
    run(3840, 2160, 3, 5, 32)
    run(1920, 1080, 3, 5, 64)
    run(8192, 8192, 1, 6, 2)
