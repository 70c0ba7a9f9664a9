#!/usr/bin/env python
"""Side measurements for the BASELINE.json configs that are not the headline bench line:
config 1 (one 512x512 frame), config 3 (one 3840x2160 frame: latency), config 4 (1080p batch through -ppm).
Kernel time from CUDA events inside the library; every output is checked against the reference decoder's digests."""
import hashlib, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W
from webp_decoder_b200 import parse as P

dg = json.loads((ROOT / "bench_data" / "digests.json").read_text())
names = sorted(dg)
pf = P.parse_batch([(ROOT / "bench_data" / n).read_bytes() for n in names], pinned=True)
idx = {n: i for i, n in enumerate(names)}
ctx = W.Context(0)
out = {}

def timed(kfs, frs, filtered=True, reps=20):
    b = ctx.upload(kfs, frs)
    for _ in range(3):
        ctx.run(b, filtered, W.TIGHT)
    ctx.kernel_time()
    for _ in range(reps):
        ctx.run(b, filtered, W.TIGHT)
    ms, n = ctx.kernel_time()
    buf, offs, sizes = ctx.download_i420(b)
    return ms / n, b, buf, offs, sizes

for label, name in (("config1_512x512_noise", "noise_512x512_q75.webp"), ("config3_4k_checker", "checker_3840x2160_q75.webp"),
                    ("config3_4k_rgbgrad", "rgbgrad_3840x2160_q75.webp")):
    for kern in (3, 2):
        ctx.set_kernel(kern)
        ctx.set_cluster(1)  # one CTA: the cluster numbers are in the last section
        i = idx[name]
        ms, b, buf, offs, sizes = timed([pf.kfs[i]], [pf.frames[i]])
        ok = hashlib.sha256(buf[int(offs[0]):int(offs[0]) + int(sizes[0])]).hexdigest() == dg[name]["yuvf"]
        w, h = dg[name]["width"], dg[name]["height"]
        out[f"{label}_kernel{kern}"] = {"latency_us": ms * 1e3, "mpixel_per_s": w * h / ms / 1e3, "launch": ctx.last_launch_config(), "bit_exact": ok}
        b.free()
ctx.set_kernel(3)
ctx.set_cluster(0)
# config 4: 1080p batch through -ppm (fused recon+filter, then the RGB kernel)
sel = [idx[n] for n in names if "1920x1080" in n]
order = [sel[k % len(sel)] for k in range(1024)]
kfs, frs = [pf.kfs[i] for i in order], [pf.frames[i] for i in order]
b = ctx.upload(kfs, frs)
import ctypes as C
L = W.load_library()
ev = []
import torch
s = torch.cuda.Stream()
ctx2 = W.Context(0, s.cuda_stream)
b2 = ctx2.upload(kfs, frs)
with torch.cuda.stream(s):
    for _ in range(2):
        ctx2.run(b2, True, W.TIGHT); ctx2.rgb(b2)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 5
    t_w = t_r = 0.0
    for _ in range(reps):
        e0.record(s); ctx2.run(b2, True, W.TIGHT); e1.record(s); ctx2.rgb(b2); e2.record(s)
        torch.cuda.synchronize()
        t_w += e0.elapsed_time(e1); t_r += e1.elapsed_time(e2)
buf, offs, sizes = ctx2.download_ppm(b2)
ok = all(hashlib.sha256(buf[int(offs[k]):int(offs[k]) + int(sizes[k])]).hexdigest() == dg[names[order[k]]]["ppm"] for k in (0, 1, 2, 3, 1023))
px = 1024 * 1920 * 1080
out["config4_ppm_1024x1080p"] = {"wavefront_ms": t_w / reps, "rgb_ms": t_r / reps, "mpixel_per_s": px / ((t_w + t_r) / reps) / 1e3,
                                  "rgb_kernel_GBps": 1024 * (1920 * 1080 * 4.5) / (t_r / reps) / 1e6, "bit_exact": ok}
print(json.dumps(out, indent=1))

# config 3 again with the image spread over a thread-block cluster (pair kernel, 16 warps per CTA)
ctx3 = W.Context(0)
lat = {}
for kern, name in ((3, "checker_3840x2160_q75.webp"), (3, "rgbgrad_3840x2160_q75.webp"), (2, "checker_3840x2160_q75.webp")):
    ctx3.set_kernel(kern)
    for cl in (1, 2, 4, 8):
        ctx3.set_cluster(cl)
        i = idx[name]
        b = ctx3.upload([pf.kfs[i]], [pf.frames[i]])
        for _ in range(3):
            ctx3.run(b, True, W.TIGHT)
        ctx3.kernel_time()
        for _ in range(20):
            ctx3.run(b, True, W.TIGHT)
        ms, n = ctx3.kernel_time()
        buf, offs, sizes = ctx3.download_i420(b)
        ok = hashlib.sha256(buf[int(offs[0]):int(offs[0]) + int(sizes[0])]).hexdigest() == dg[name]["yuvf"]
        lat[f"kernel {kern} {name} cluster<= {cl}"] = {"latency_us": ms / n * 1e3, "launch": ctx3.last_launch_config(), "bit_exact": ok}
        b.free()
print(json.dumps({"config3_cluster": lat}, indent=1))
