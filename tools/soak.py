#!/usr/bin/env python
"""Randomised soak of the wavefront kernels against the CPU oracle: random batch sizes, frame sizes, content density,
launch shapes (warps per image, images per CTA, cluster size) and kernel generations, for a given number of seconds.
Development aid, run under gpurun:  python tools/soak.py [seconds] [seed] [--clusters]
--clusters: only the shapes of the cluster kernels (vp8_mb_split / vp8_mb_pairs<16> on 2, 4, 8 CTAs per image): few big frames,
each oracle frame computed once and decoded under every cluster size and flavour, repeatedly (the hand-over between
warps / CTAs is timing dependent: the same input many times is the race check)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import webp_decoder_b200 as W  # noqa: E402
from vp8fix import Oracle, fuzz_frame  # noqa: E402


def soak(ctx, orc, rng, seconds=None, rounds=None):
    """Runs until `seconds` have passed or `rounds` rounds are done; raises AssertionError on the first mismatch.
    Returns (rounds, frames)."""
    t0, done, frames_done = time.time(), 0, 0
    while (seconds is None or time.time() - t0 < seconds) and (rounds is None or done < rounds):
        n = int(rng.choice([1, 2, 5, 13, 40, 150, 300]))
        big = rng.random() < 0.2
        frames = []
        for i in range(n):
            w = int(rng.integers(1, 700 if big and n <= 13 else 130))
            h = int(rng.integers(1, 1300 if big and n <= 13 else 130))
            frames.append(fuzz_frame(int(rng.integers(1 << 30)), w, h, density=float(rng.choice([0.0, 0.02, 0.3, 0.9])),
                                     amp=int(rng.choice([5, 40, 400, 2500])), raw=bool(rng.integers(2))))
        kernel = int(rng.choice([2, 3, 3, 3]))
        warps = int(rng.choice([0, 0, 4, 8, 16]))
        per_sm = int(rng.choice([0, 0, 1, 2, 3, 5, 7]))
        cluster = int(rng.choice([0, 0, 1, 2, 4, 8]))
        filtered = bool(rng.integers(2))
        ctx.set_kernel(kernel)
        ctx.set_tuning(warps, per_sm)
        split = bool(rng.integers(2))  # clusters of 4 / 8 CTAs, fused mode: vp8_mb_split or the one-warp-per-row-pair kernel
        ctx.set_cluster(cluster, split)
        kfs, ds = [f.header() for f in frames], [f.cstruct() for f in frames]
        want = [orc.decode_i420(f, filtered) for f in frames]
        api = str(rng.choice(["batch", "batch", "pipelined", "ppm"]))
        if api == "batch":
            outs = ctx.decode_i420(kfs, ds, filtered=filtered)
        elif api == "pipelined":  # chunked, compact or dense transport, several host threads
            ctx.set_transport([True, False, "auto"][int(rng.integers(3))], int(rng.choice([1, 3, 8])))
            buf = np.empty(ctx.decode_bytes(kfs), np.uint8)
            offs, szs = ctx.decode_into(kfs, ds, buf, filtered=filtered, chunk=int(rng.choice([1, 3, 16, 64])))
            outs = [buf[int(o):int(o) + int(z)] for o, z in zip(offs, szs)]
        else:  # m08 on top of the filtered frames
            filtered = True
            want = [np.frombuffer(orc.ppm(orc.rgb(orc.decode_i420(f, True), f.width, f.height), f.width, f.height), np.uint8) for f in frames]
            outs = [np.frombuffer(p, np.uint8) for p in ctx.decode_ppm(kfs, ds)]
        cfg = ctx.last_launch_config()
        bad = [i for i, (w, o) in enumerate(zip(want, outs)) if not np.array_equal(o, w)]
        assert not bad, (f"MISMATCH round {done} ({api}): kernel {kernel} warps {warps} per_sm {per_sm} cluster {cluster} split {split} filtered {filtered} "
                         f"launch {cfg}: frames {bad[:8]} of {n}, e.g. {frames[bad[0]].width}x{frames[bad[0]].height}")
        done += 1
        frames_done += n
    return done, frames_done


def soak_clusters(ctx, orc, rng, seconds):
    t0, done, frames_done = time.time(), 0, 0
    while time.time() - t0 < seconds:
        n = int(rng.choice([1, 1, 2, 3, 6]))
        frames = [fuzz_frame(int(rng.integers(1 << 30)), int(rng.integers(17, 2100)), int(rng.integers(530, 2300)),
                             density=float(rng.choice([0.0, 0.05, 0.3, 0.9])), amp=int(rng.choice([5, 40, 400, 2500])),
                             raw=bool(rng.integers(2))) for _ in range(n)]
        kfs, ds = [f.header() for f in frames], [f.cstruct() for f in frames]
        want = {flt: [orc.decode_i420(f, flt) for f in frames] for flt in (False, True)}
        for rep in range(6):
            for cluster in (2, 4, 8):
                for split in (False, True):
                    for flt in (False, True):
                        ctx.set_cluster(cluster, split)
                        outs = ctx.decode_i420(kfs, ds, filtered=flt)
                        cfg = ctx.last_launch_config()
                        bad = [i for i, (w, o) in enumerate(zip(want[flt], outs)) if not np.array_equal(o, w)]
                        assert not bad, (f"MISMATCH round {done} rep {rep}: cluster {cluster} split {split} filtered {flt} launch {cfg}: "
                                         f"frames {bad} of {n}, e.g. {frames[bad[0]].width}x{frames[bad[0]].height}")
                        frames_done += n
        done += 1
    ctx.set_cluster(0, True)
    return done, frames_done


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    seconds = float(args[0]) if len(args) > 0 else 60.0
    rng = np.random.default_rng(int(args[1]) if len(args) > 1 else 1)
    t0 = time.time()
    try:
        if "--clusters" in sys.argv:
            rounds, frames_done = soak_clusters(W.Context(0), Oracle(), rng, seconds)
        else:
            rounds, frames_done = soak(W.Context(0), Oracle(), rng, seconds=seconds)
    except AssertionError as e:
        print(e)
        return 1
    print(f"soak ok: {rounds} rounds, {frames_done} frames, {time.time() - t0:.0f} s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
