#!/usr/bin/env python
"""CPU arm of the benchmark: the reference's own m06+m07(+m08) on the host cores, stage-matched with the GPU number.

TEST/MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `bench.py --impl reference`).

Each worker process loads the UNMODIFIED reference (oracle/_ref/libref_decode.so, built from /root/reference by
oracle/Makefile), decodes the tokens of every input once with the reference's own m01..m05 (untimed, as the GPU
arm's inputs are also already parsed), then times `vp8_reconstruct_keyframe_yuv_filtered` (or `_yuv`, or
`..._filtered` + `yuv420_write_ppm_fd` to /dev/null) over its share of a bounded number of frames.
One process per core; throughput = pixels of all workers / slowest worker's time.
If oracle/_ref is absent the oracle port (oracle/liboracle.so) is timed instead and `kind` says "port".
Prints one JSON object.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def _worker(args):
    files, mode, reps, kind = args
    import numpy as np  # noqa: F401
    from vp8fix import Oracle, Reference
    if kind == "reference":
        ref = Reference()
        frames = [ref.parse_webp(Path(f).read_bytes()) for f in files]
        L = ref.lib
        from vp8fix import Yuv420Image
        fn = L.vp8_reconstruct_keyframe_yuv if mode == "yuv" else L.vp8_reconstruct_keyframe_yuv_filtered
        structs = [(fr.header(), fr.cstruct()) for fr in frames]
        devnull = os.open(os.devnull, os.O_WRONLY)
        img = Yuv420Image()

        def one(i):
            h, d = structs[i]
            assert fn(C.byref(h), C.byref(d), C.byref(img)) == 0
            if mode == "ppm":
                assert L.yuv420_write_ppm_fd(devnull, C.byref(img)) == 0
            L.yuv420_free(C.byref(img))
    else:
        # port: frames still come from the reference-free product parser (input provider only)
        sys.path.insert(0, str(ROOT))
        from webp_decoder_b200 import parse as P
        from test_oracle import ParsedAsFrame
        orc = Oracle()
        pf = P.parse_batch([Path(f).read_bytes() for f in files], threads=1)
        frames = [ParsedAsFrame(pf.kfs[i], pf.frames[i]) for i in range(len(files))]

        def one(i):
            out = orc.decode_i420(frames[i], mode != "yuv")
            if mode == "ppm":
                orc.rgb(out, frames[i].width, frames[i].height)
    one(0)  # warm the caches and the page tables
    t0 = time.perf_counter()
    px = 0
    for _ in range(reps):
        for i, fr in enumerate(frames):
            one(i)
            px += fr.width * fr.height
    return px, time.perf_counter() - t0, reps * len(frames)


def run(files, mode="yuvf", procs=None, seconds=12.0):
    from vp8fix import Reference
    kind = "reference" if Reference.available() else "port"
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        # calibrate with one repetition on every core at once (contention included), then size the sample
        cal = pool.map(_worker, [(files, mode, 1, kind)] * procs)
        per_rep = max(t for _, t, _ in cal)
        reps = max(1, int(seconds / max(per_rep, 1e-3)))
        res = pool.map(_worker, [(files, mode, reps, kind)] * procs)
    px = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    frames = sum(r[2] for r in res)
    return {
        "value": px / slowest / 1e6, "unit": "Mpixel/s", "cores": procs, "kind": kind,
        "frames_per_s": frames / slowest,
        "sample": f"{frames} frames ({reps} passes over {len(files)} distinct inputs on each of {procs} processes), "
                  f"{slowest:.1f} s, mode -{mode}, token decode excluded",
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--mode", default="yuvf", choices=["yuv", "yuvf", "ppm"])
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=12.0)
    a = ap.parse_args()
    print(json.dumps(run(a.files, a.mode, a.procs or None, a.seconds)))
