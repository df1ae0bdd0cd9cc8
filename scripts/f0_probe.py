"""Development probe: level times and packed-path statistics of the encode stage (run with RIC_FWD0=0/1)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from rududu_image_codec_b200 import capi
if os.environ.get("RIC_LIB"):  # A/B timing of a variant build (scripts/build_variant.sh)
    capi.LIB_PATH = os.path.abspath(os.environ["RIC_LIB"])
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
    runs = []
    for _ in range(9):  # median of 9 calls per level: single calls scatter by about 2 %
        c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
        torch.cuda.synchronize()
        runs.append(c.level_times(0))
    med = [sorted(r[k] for r in runs)[len(runs) // 2] for k in range(len(runs[0]))]
    print("%dx%dx%d n=%d q=%d FWD0=%s: level ms (median of 9) %s  level-0 min %.3f max %.3f" % (
          w, h, ch, n, q, os.environ.get("RIC_FWD0", "0"), ["%.3f" % t for t in med],
          min(r[0] for r in runs), max(r[0] for r in runs)), flush=True)
    c.close()

if __name__ == "__main__":
    run(3840, 2160, 3, 5, 32)
    run(1920, 1080, 3, 5, 64)
    run(8192, 8192, 1, 6, 2)
