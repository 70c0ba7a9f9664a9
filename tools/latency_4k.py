#!/usr/bin/env python
"""BASELINE config 3: device time of recon + loop filter of ONE 3840x2160 frame, by cluster size (CTAs per image); also one
1080p frame and small batches of either (the shapes a chunk of the pipelined calls launches), and the unfiltered flavour
(the reconstruction chain alone). Development aid, run under gpurun: python tools/latency_4k.py [--quick]
(VP8_GPU_LIB selects a library variant, VP8_GPU_SPLIT=0/1 the cluster kernel)"""
import hashlib, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W
from webp_decoder_b200 import parse as P

quick = "--quick" in sys.argv
lf0 = "--lf0" in sys.argv  # loop-filter level forced to 0 (bytes then differ from the digests): what the filter ARITHMETIC costs the chain
dg = json.loads((ROOT / "bench_data" / "digests.json").read_text())
ctx = W.Context(0)
cases = [("checker_3840x2160_q75.webp", 1), ("rgbgrad_3840x2160_q75.webp", 1), ("rgbgrad_1920x1080_q75.webp", 1),
         ("noise_1920x1080_q75.webp", 1), ("rgbgrad_3840x2160_q75.webp", 18), ("rgbgrad_1920x1080_q75.webp", 37),
         ("rgbgrad_1920x1080_q75.webp", 64)]
for name, copies in cases:
    pf = P.parse_batch([(ROOT / "bench_data" / name).read_bytes()], pinned=True)
    if lf0:
        pf.frames[0].lf_level = 0
    b = ctx.upload([pf.kfs[0]] * copies, [pf.frames[0]] * copies)
    for cl in ((0,) if quick or copies > 1 else (1, 2, 4, 8, 0)):
        for filtered in (True, False):
            if not filtered and (cl or copies > 1):
                continue
            ctx.set_cluster(cl)
            for _ in range(3):
                ctx.run(b, filtered, W.TIGHT)
            ctx.kernel_time()
            for _ in range(20):
                ctx.run(b, filtered, W.TIGHT)
            ms, n = ctx.kernel_time()
            buf, offs, sizes = ctx.download_i420(b)
            want = dg[name]["yuvf" if filtered else "yuv"]
            ok = all(hashlib.sha256(buf[int(o):int(o) + int(s)]).hexdigest() == want for o, s in zip(offs, sizes))
            print(f"{name} x{copies} cluster {cl} {'yuvf' if filtered else 'yuv '}: {ms / n * 1e3:8.1f} us  bit_exact {ok}  "
                  f"{ctx.last_launch_config()}", flush=True)
    b.free()
    pf.free()
