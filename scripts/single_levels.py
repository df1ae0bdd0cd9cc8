"""Per-level launch times for single images (development aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from rududu_image_codec_b200 import capi
from rududu_image_codec_b200.synth import synth_image
for (w, h, ch, lv) in [(3840, 2160, 3, 5), (8192, 8192, 1, 6)]:
    c = capi.Context(w, h, ch, lv, max_batch=1)
    pitch = (w + 7) & ~7
    src = torch.zeros((ch, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :w] = torch.from_numpy(synth_image(0, w, h, ch)).cuda()
    ar = torch.zeros(c.image_arena_bytes + 64, dtype=torch.uint8, device="cuda")
    dst = torch.zeros((ch, h, pitch), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    c.set_profiling(True)
    for _ in range(5):
        c.encode_u8_device(src.data_ptr(), pitch, 1, 9, ar.data_ptr(), st)
        c.decode_u8_device(ar.data_ptr(), 1, 9, dst.data_ptr(), pitch, st)
    torch.cuda.synchronize()
    print(w, h, ch, "enc", [round(x * 1e3, 1) for x in c.level_times(0)], "dec", [round(x * 1e3, 1) for x in c.level_times(1)], "us")
    c.close()
