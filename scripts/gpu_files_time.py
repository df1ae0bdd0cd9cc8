"""Time ric_compress_u8_gpu / ric_decompress_u8_gpu on n 1080p RGB images (pinned host buffers).
usage: [RIC_TRACE=1] python scripts/gpu_files_time.py [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rududu_image_codec_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = capi.lib()
for rep in range(2):
    t0 = time.perf_counter()
    r = bench.ric_file_throughput_device(capi, L, 0, n, 9)
    print("%.1f s total" % (time.perf_counter() - t0), r, flush=True)
