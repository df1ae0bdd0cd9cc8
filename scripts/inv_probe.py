"""Development probe: level times of the decode stage."""
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
    pitch = (w + 15) & ~7
    src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :, :w] = torch.from_numpy(imgs).cuda().repeat((n + 3) // 4, 1, 1, 1)[:n]
    ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
    dst = torch.zeros_like(src)
    st = torch.cuda.current_stream().cuda_stream
    c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    torch.cuda.synchronize()
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    bench.unfold_arenas_(c, ar, min(n, 4))  # signed coefficients, as the entropy decoder leaves them
    per = c.image_arena_bytes
    for i in range(4, n):
        ar[i * per:(i + 1) * per].copy_(ar[(i % 4) * per:((i % 4) + 1) * per])
    c.set_profiling(True)
    for _ in range(4):
        c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
    torch.cuda.synchronize()
    c.path_stats()
    runs = []
    for _ in range(9):  # median of 9 calls per level
        c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
        torch.cuda.synchronize()
        runs.append(c.level_times(1))
    lt = [sorted(r[k] for r in runs)[len(runs) // 2] for k in range(len(runs[0]))]
    print("%dx%dx%d n=%d q=%d INV0=%s decode level ms (median of 9) %s stats %s" % (w, h, ch, n, q, os.environ.get("RIC_INV0", "0"), ["%.3f" % t for t in lt], c.path_stats()), flush=True)
    c.close()

if __name__ == "__main__":
    run(3840, 2160, 3, 5, 32)
    run(1920, 1080, 3, 5, 64)
    run(8192, 8192, 1, 6, 2)
