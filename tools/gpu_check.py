#!/usr/bin/env python
"""GPU bring-up diagnostic: runs targeted frames through libvp8gpu.so and reports WHERE the output first
differs from the CPU oracle (plane, pixel, macroblock, sub-block). Development aid, run under gpurun:

    python tools/gpu_check.py [--quick]
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import webp_decoder_b200 as W  # noqa: E402
from vp8fix import Oracle, fuzz_frame  # noqa: E402


def locate(idx, w, h):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if idx < w * h:
        pl, x, y, mbw = "Y", idx % w, idx // w, 16
    elif idx < w * h + cw * ch:
        j = idx - w * h
        pl, x, y, mbw = "U", j % cw, j // cw, 8
    else:
        j = idx - w * h - cw * ch
        pl, x, y, mbw = "V", j % cw, j // cw, 8
    return f"{pl}({x},{y}) mb=({x // mbw},{y // mbw}) in-mb=({x % mbw},{y % mbw})"


def report(name, got, want, w, h, limit=6):
    if got.shape == want.shape and np.array_equal(got, want):
        print(f"  ok    {name}")
        return True
    if got.shape != want.shape:
        print(f"  FAIL  {name}: size {got.shape} vs {want.shape}")
        return False
    bad = np.flatnonzero(got != want)
    print(f"  FAIL  {name}: {bad.size} of {want.size} bytes differ")
    for i in bad[:limit]:
        print(f"          {locate(int(i), w, h)} got {got[i]} want {want[i]}")
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--kernel", type=int, default=3, help="2 = vp8_mb_pairs for every batch size, 3 = lockstep flavour for big batches")
    args = ap.parse_args()
    orc = Oracle()
    ctx = W.Context(0)
    ctx.set_kernel(args.kernel)
    nofilt = dict(lf_level=0, segmentation_enabled=0, lf_delta_enabled=0)
    plain = dict(segmentation_enabled=0, lf_delta_enabled=0, lf_sharpness=0, lf_use_simple=0)
    cases = [
        ("1mb i16 no-coeff", fuzz_frame(1, 16, 16, density=0.0, bpred_frac=0.0, **nofilt)),
        ("1mb i16 coeff", fuzz_frame(2, 16, 16, density=0.3, bpred_frac=0.0, **nofilt)),
        ("1mb bpred no-coeff", fuzz_frame(3, 16, 16, density=0.0, bpred_frac=1.0, **nofilt)),
        ("1mb bpred coeff", fuzz_frame(4, 16, 16, density=0.3, bpred_frac=1.0, **nofilt)),
        ("row of 4 i16", fuzz_frame(5, 64, 16, density=0.3, bpred_frac=0.0, **nofilt)),
        ("row of 4 bpred", fuzz_frame(6, 64, 16, density=0.3, bpred_frac=1.0, **nofilt)),
        ("col of 4 i16", fuzz_frame(7, 16, 64, density=0.3, bpred_frac=0.0, **nofilt)),
        ("col of 4 bpred", fuzz_frame(8, 16, 64, density=0.3, bpred_frac=1.0, **nofilt)),
        ("4x4 mixed", fuzz_frame(9, 64, 64, density=0.3, **nofilt)),
        ("4x4 mixed crop 50x37", fuzz_frame(10, 50, 37, density=0.3, **nofilt)),
        ("normal lf 2x2 level 20", fuzz_frame(11, 32, 32, density=0.3, lf_level=20, **plain)),
        ("normal lf 4x4 level 50", fuzz_frame(12, 64, 64, density=0.3, lf_level=50, **plain)),
        ("simple lf 4x4", fuzz_frame(13, 64, 64, density=0.3, lf_level=30, segmentation_enabled=0, lf_delta_enabled=0,
                                     lf_sharpness=0, lf_use_simple=1)),
        ("all knobs 5x3 crop", fuzz_frame(14, 77, 45, density=0.3)),
        ("overflow amp", fuzz_frame(15, 64, 48, density=0.5, amp=2500)),
        ("raw slots", fuzz_frame(16, 64, 48, density=0.4, raw=True)),
        ("wide 40x3", fuzz_frame(17, 640, 48, density=0.2)),
        ("tall 3x40", fuzz_frame(18, 48, 640, density=0.2)),
    ]
    if not args.quick:
        cases += [(f"fuzz {s}", fuzz_frame(100 + s, int(np.random.default_rng(s).integers(1, 300)),
                                           int(np.random.default_rng(s + 999).integers(1, 300)),
                                           amp=[5, 40, 400, 2500][s % 4], density=[0.05, 0.2, 0.6][s % 3], raw=bool(s & 1)))
                  for s in range(40)]
    ok = True
    for name, fr in cases:
        w, h = fr.width, fr.height
        print(f"[{name}] {w}x{h} lf_level={fr.params['lf_level']} simple={fr.params['lf_use_simple']} seg={fr.params['segmentation_enabled']}")
        kf, d = fr.header(), fr.cstruct()
        for warps in ((4,) if args.quick else (4, 8, 16)):
            ctx.set_tuning(warps, 0)
            want_u = orc.decode_i420(fr, False)
            got_u = ctx.decode_i420([kf], [d], filtered=False)[0]
            ok &= report(f"recon        (warps={warps})", got_u, want_u, w, h)
            want_f = orc.decode_i420(fr, True)
            got_f = ctx.decode_i420([kf], [d], filtered=True)[0]
            ok &= report(f"recon+filter (warps={warps})", got_f, want_f, w, h)
        ctx.set_tuning(0, 0)
        # staged path: recon -> padded, filter in place, rgb
        b = ctx.recon([kf], [d])
        y, u, v = ctx.download_padded(b, 0)
        oy, ou, ov = orc.recon_padded(fr)
        ok &= report("stage recon padded Y", y.ravel(), oy.ravel(), y.shape[1], y.shape[0])
        ctx.filter(b)
        y, u, v = ctx.download_padded(b, 0)
        orc.loopfilter_padded(fr, oy, ou, ov)
        pw, ph = y.shape[1], y.shape[0]
        ok &= report("stage filter padded", np.concatenate([y.ravel(), u.ravel(), v.ravel()]),
                     np.concatenate([oy.ravel(), ou.ravel(), ov.ravel()]), pw, ph)
        ctx.rgb(b)
        buf, offs, sizes = ctx.download_ppm(b)
        ppm = buf[int(offs[0]):int(offs[0]) + int(sizes[0])].tobytes()
        want_ppm = orc.ppm(orc.rgb(want_f, w, h), w, h)
        if ppm == want_ppm:
            print("  ok    ppm")
        else:
            ok = False
            hl = len(want_ppm) - w * h * 3
            g = np.frombuffer(ppm[hl:], np.uint8)
            e = np.frombuffer(want_ppm[hl:], np.uint8)
            bad = np.flatnonzero(g != e) if g.shape == e.shape else np.array([0])
            print(f"  FAIL  ppm header_ok={ppm[:hl] == want_ppm[:hl]} {bad.size} rgb bytes differ; first at px {bad[:5] // 3}")
        b.free()
    # batch of many frames at once
    frames = [fuzz_frame(500 + s, 64 + 16 * (s % 5), 48 + 16 * (s % 3), density=0.2) for s in range(64)]
    t0 = time.time()
    outs = ctx.decode_i420([f.header() for f in frames], [f.cstruct() for f in frames], filtered=True)
    good = sum(np.array_equal(o, orc.decode_i420(f, True)) for o, f in zip(outs, frames))
    print(f"[batch of 64 mixed sizes] {good}/64 frames equal ({time.time() - t0:.2f}s)")
    ok &= good == 64
    if args.kernel == 3:
        # lockstep flavour: several images per CTA, forced through the tuning knobs so that a small batch takes it
        frames = [fuzz_frame(700 + s, [64, 129, 300, 17, 48][s % 5], [48, 129, 90, 33, 200][s % 5],
                             density=[0.3, 0.02, 0.9, 0.0][s % 4], amp=[40, 400, 2500][s % 3], raw=bool(s & 1)) for s in range(45)]
        want_f = [orc.decode_i420(f, True) for f in frames]
        want_u = [orc.decode_i420(f, False) for f in frames]
        for per_cta in (2, 3, 7):
            ctx.set_tuning(4, per_cta)
            for filtered, want in ((True, want_f), (False, want_u)):
                outs = ctx.decode_i420([f.header() for f in frames], [f.cstruct() for f in frames], filtered=filtered)
                good = sum(np.array_equal(o, w) for o, w in zip(outs, want))
                cfg = ctx.last_launch_config()
                print(f"[lockstep {per_cta} images per CTA, filtered={filtered}] {good}/{len(frames)} frames equal, launch {cfg}")
                ok &= good == len(frames) and cfg["images_per_cta"] == per_cta
        ctx.set_tuning(0, 0)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
