#!/usr/bin/env python
"""BASELINE config 3: device time of recon + loop filter of ONE 3840x2160 frame, by cluster size (CTAs per image).
Development aid, run under gpurun: python tools/latency_4k.py   (VP8_GPU_LIB selects a library variant)"""
import hashlib, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W
from webp_decoder_b200 import parse as P

dg = json.loads((ROOT / "bench_data" / "digests.json").read_text())
ctx = W.Context(0)
for name in ("checker_3840x2160_q75.webp", "rgbgrad_3840x2160_q75.webp"):
    pf = P.parse_batch([(ROOT / "bench_data" / name).read_bytes()], pinned=True)
    b = ctx.upload([pf.kfs[0]], [pf.frames[0]])
    for cl in (1, 2, 4, 8, 0):
        ctx.set_cluster(cl)
        for _ in range(3):
            ctx.run(b, True, W.TIGHT)
        ctx.kernel_time()
        for _ in range(20):
            ctx.run(b, True, W.TIGHT)
        ms, n = ctx.kernel_time()
        buf, offs, sizes = ctx.download_i420(b)
        ok = hashlib.sha256(buf[int(offs[0]):int(offs[0]) + int(sizes[0])]).hexdigest() == dg[name]["yuvf"]
        print(f"{name} cluster {cl}: {ms / n * 1e3:8.1f} us  bit_exact {ok}  {ctx.last_launch_config()}")
    b.free()
    pf.free()
