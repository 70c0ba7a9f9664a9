#!/usr/bin/env python
"""Geometry sweep of the m09 kernels on the GPU: random picture sizes (1 .. 4000 wide, scanlines around the 16-byte segment and
65535-byte stored-block boundaries) as random I420 through the drop-in yuv420_write_png_fd (RGB kernel + vp8_png_frame /
vp8_png_finish) against the host framing (vp8_gpu_png_frame) of the same device RGB, plus zlib's own CRC-32 / inflate / Adler-32
of every file.   python tools/png_soak.py [rounds] [seed]"""
import ctypes as C
import struct
import sys
import zlib
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import webp_decoder_b200 as W  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
L = W.load_library()
L.vp8_gpu_png_bound.argtypes, L.vp8_gpu_png_bound.restype = [C.c_uint32, C.c_uint32], C.c_size_t
L.vp8_gpu_png_frame.argtypes, L.vp8_gpu_png_frame.restype = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p], C.c_size_t
special_w = [1, 2, 3, 4, 5, 6, 10, 21, 85, 341, 1365, 5461, 21845, 16383, 16, 1920, 3840]
total = 0
for k in range(rounds):
    if k < len(special_w):
        w, h = special_w[k], int(rng.integers(1, 40))
    elif k % 3 == 0:
        w, h = int(rng.integers(1, 64)), int(rng.integers(1, 3000))
    else:
        w, h = int(rng.integers(1, 4000)), int(rng.integers(1, 1200))
    cw, ch = (w + 1) // 2, (h + 1) // 2
    i420 = rng.integers(0, 256, w * h + 2 * cw * ch, dtype=np.uint8)
    png = W.yuv420_write_png(i420, w, h)
    ppm = W.yuv420_write_ppm(i420, w, h)
    rgb = np.frombuffer(ppm, np.uint8)[len(ppm) - 3 * w * h:].copy()
    host = np.empty(L.vp8_gpu_png_bound(w, h), np.uint8)
    n = L.vp8_gpu_png_frame(rgb.ctypes.data, w, h, host.ctypes.data)
    assert png == host[:n].tobytes(), (k, w, h)
    n_idat, = struct.unpack(">I", png[33:37])
    assert zlib.crc32(png[37:41 + n_idat]) == struct.unpack(">I", png[41 + n_idat:45 + n_idat])[0], (k, w, h)
    assert zlib.decompress(png[41:41 + n_idat]) == np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, 3 * w)], axis=1).tobytes(), (k, w, h)
    total += len(png)
print(f"png soak ok: {rounds} pictures, {total / 1e6:.1f} MB of files")
