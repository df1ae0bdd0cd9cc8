"""Small fixed workload for ncu captures: N 1080p RGB images, encode + decode, device resident."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w, h, ch, q = 1920, 1080, 3, 9
img = synth_image(0, w, h, ch)
c = capi.Context(w, h, ch, 5, max_batch=n)
pitch = (w + 15) & ~7
src = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
src[:, :, :, :w] = torch.from_numpy(img).cuda()[None]
ar = torch.zeros(n * c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
dst = torch.zeros((n, ch, h, pitch), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    c.encode_u8_device(src.data_ptr(), pitch, n, q, ar.data_ptr(), st)
    c.decode_u8_device(ar.data_ptr(), n, q, dst.data_ptr(), pitch, st)
torch.cuda.synchronize()
print("ok")
