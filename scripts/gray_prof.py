"""Development aid: a few encode-stage calls on 8192x8192 gray (BASELINE configs[2] shape), for ncu."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image
w, h, ch, levels, n, q = 8192, 8192, 1, 6, 2, 9
img = synth_image(0, w, h, ch)
c = capi.Context(w, h, ch, levels, max_batch=n)
src = torch.from_numpy(img).cuda()[None].repeat(n, 1, 1, 1).contiguous()
ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    c.encode_u8_device(src.data_ptr(), w, n, q, ar.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
c.close()
