"""Development probe: single-image and batch stage times with and without programmatic dependent launch (RIC_PDL=0/1)."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import importlib.util
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
import torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image
dev = torch.device("cuda:0")
out = {"pdl": os.environ.get("RIC_PDL", "1")}
for (w, h, ch, lv) in ((3840, 2160, 3, 5), (8192, 8192, 1, 6), (1920, 1080, 3, 5)):
    r = bench.single_image_latency(capi, synth_image, dev, w, h, ch, lv, 9)
    out["%dx%dx%d" % (w, h, ch)] = (round(r["encode_ms"], 4), round(r["decode_ms"], 4))
# batch: 32 x 4K through the device calls, profiling off (PDL active), event-timed over 10 calls
w, h, ch, n, q = 3840, 2160, 3, 32, 9
c = capi.Context(w, h, ch, 5, max_batch=n)
pitch = (w + 15) & ~15
import numpy as np
imgs = np.stack([synth_image(i, w, h, ch) for i in range(4)])
src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
src[:, :, :, :w] = torch.from_numpy(imgs).cuda().repeat(8, 1, 1, 1)
ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
dst = torch.zeros_like(src)
st = torch.cuda.current_stream().cuda_stream
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
torch.cuda.synchronize()
te = td = 0.0
for it in range(10):
    e[0].record(); c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    e[1].record(); c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
    e[2].record(); torch.cuda.synchronize()
    te += e[0].elapsed_time(e[1]); td += e[1].elapsed_time(e[2])
out["batch32_4k_ms"] = (round(te / 10, 4), round(td / 10, 4))
c.close()
print(json.dumps(out))
