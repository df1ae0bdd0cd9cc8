"""Development aid: a few encode-stage calls on the bench workload, small enough to run under ncu."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from rududu_image_codec_b200 import capi
if os.environ.get("RIC_LIB"):  # A/B profiling of a variant build (scripts/build_variant.sh)
    capi.LIB_PATH = os.path.abspath(os.environ["RIC_LIB"])
from rududu_image_codec_b200.synth import synth_image
w, h, ch, levels, n, q = 3840, 2160, 3, 5, int(os.environ.get("N", "16")), 9
imgs = np.stack([synth_image(i, w, h, ch) for i in range(4)])
c = capi.Context(w, h, ch, levels, max_batch=n)
pitch = (w + 15) & ~7
src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
src[:, :, :, :w] = torch.from_numpy(imgs).cuda().repeat((n + 3) // 4, 1, 1, 1)[:n]
ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
dst = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
import importlib.util
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
torch.cuda.synchronize()
dec = ar.clone()
bench.unfold_arenas_(c, dec, 4)  # signed coefficients, as the entropy decoder leaves them
per = c.image_arena_bytes
for i in range(4, n):
    dec[i * per:(i + 1) * per].copy_(dec[(i % 4) * per:((i % 4) + 1) * per])
for _ in range(2):
    c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    c.decode_u8_device(dec.data_ptr(), n, q, dst.data_ptr(), pitch, st)
torch.cuda.synchronize()
print("ok")
c.close()
