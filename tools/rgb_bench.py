#!/usr/bin/env python
"""m08 alone: 1024 x 1080p I420 (resident) -> RGB24, CUDA-event time per launch, GB/s against the measured HBM peak, and
a byte check of every frame against the reference decoder's -ppm digests (bench_data/digests.json)."""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W
from webp_decoder_b200 import parse as P

files = ["noise_1920x1080_q75.webp", "rgbgrad_1920x1080_q75.webp", "checker_1920x1080_q75.webp", "diag_1920x1080_q75.webp"]
dg = json.loads((ROOT / "bench_data" / "digests.json").read_text())
pf = P.parse_batch([(ROOT / "bench_data" / f).read_bytes() for f in files], pinned=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
order = [i % 4 for i in range(n)]
ctx = W.Context(0)
b = ctx.upload([pf.kfs[i] for i in order], [pf.frames[i] for i in order])
ctx.run(b, True, W.TIGHT)
for _ in range(3):
    ctx.rgb(b)
ctx.rgb_time()
for _ in range(10):
    ctx.rgb(b)
ms, k = ctx.rgb_time()
buf, offs, sizes = ctx.download_ppm(b)
bad = [i for i in range(n) if hashlib.sha256(buf[int(offs[i]):int(offs[i]) + int(sizes[i])]).hexdigest() != dg[files[order[i]]]["ppm"]]
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
gb = n * 1920 * 1080 * 4.5 / 1e9
print(json.dumps({"kernel": "vp8_i420_to_rgb", "frames": n, "ms_per_launch": ms / k, "launches": k, "algorithmic_GB": gb,
                  "GBps": gb / (ms / k / 1e3), "frac_of_hbm_peak": gb / (ms / k / 1e3) / peak, "mismatching_frames": len(bad)}))
