#!/usr/bin/env python
"""CPU arm of the benchmark: the reference's own m06+m07(+m08) on the host cores, stage-matched with the GPU number.

TEST/MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `bench.py --impl reference`).

Each worker process loads the UNMODIFIED reference (oracle/_ref/libref_decode[_v3].so, built from /root/reference by
oracle/Makefile; the x86-64-v3 build when this host's CPU has AVX2/BMI2/FMA) and runs one of two loops over its share of
a bounded number of frames:
  stage  the tokens of every input are decoded once with the reference's own m01..m05 (untimed, as the GPU arm's inputs
         are also already parsed), then `vp8_reconstruct_keyframe_yuv_filtered` (or `_yuv`, or `..._filtered` +
         `yuv420_write_ppm_fd` to /dev/null) is timed: stage-matched with the GPU kernel / e2e numbers;
  whole  the reference decoder's whole in-memory path per frame, as main.c:630-702 (cmd_yuvf) runs it: container +
         header + `vp8_decode_decoded_frame` (bool decoder, tokens) + reconstruction + loop filter (+ PPM), .webp bytes in
         memory to pixels in memory: the counterpart of the GPU arm's e2e_from_webp.
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


def ref_library():
    """(path, flags) of the reference build this host should run."""
    import vp8fix
    v3 = vp8fix.REF_DIR / "libref_decode_v3.so"
    try:
        flags = set(next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split())
    except (OSError, StopIteration):
        flags = set()
    if v3.exists() and {"avx2", "bmi2", "fma"} <= flags:
        return v3, "gcc -O3 -march=x86-64-v3 (portable stand-in for -march=native: built off the box)"
    return vp8fix.REF_DIR / "libref_decode.so", "gcc -O3 (x86-64 baseline)"


def _worker(args):
    files, mode, reps, kind, whole = args
    import numpy as np  # noqa: F401
    import vp8fix
    from vp8fix import Oracle, Reference
    if kind == "reference":
        ref = Reference(ref_library()[0])
        frames = [ref.parse_webp(Path(f).read_bytes()) for f in files]
        L = ref.lib
        from vp8fix import ByteSpan, DecodedFrame, KeyFrameHeader, Yuv420Image, u8p
        fn = L.vp8_reconstruct_keyframe_yuv if mode == "yuv" else L.vp8_reconstruct_keyframe_yuv_filtered
        structs = [(fr.header(), fr.cstruct()) for fr in frames]
        devnull = os.open(os.devnull, os.O_WRONLY)
        img = Yuv420Image()
        if whole:
            spans = []
            for f in files:
                data = Path(f).read_bytes()
                size = int.from_bytes(data[16:20], "little")
                payload = (C.c_uint8 * size).from_buffer_copy(data[20:20 + size])
                spans.append((payload, ByteSpan(C.cast(payload, u8p), size)))

            def one(i):
                kf, d = KeyFrameHeader(), DecodedFrame()
                span = spans[i][1]
                assert L.vp8_parse_keyframe_header(span, C.byref(kf)) == 0
                assert L.vp8_decode_decoded_frame(span, C.byref(d)) == 0
                assert fn(C.byref(kf), C.byref(d), C.byref(img)) == 0
                if mode == "ppm":
                    assert L.yuv420_write_ppm_fd(devnull, C.byref(img)) == 0
                L.yuv420_free(C.byref(img))
                L.vp8_decoded_frame_free(C.byref(d))
        else:
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

        datas = [Path(f).read_bytes() for f in files]

        def one(i):
            fr = frames[i]
            if whole:  # the product parser stands in for m01..m05 when the reference is absent
                one_pf = P.parse_batch([datas[i]], threads=1)
                fr = ParsedAsFrame(one_pf.kfs[0], one_pf.frames[0])
            out = orc.decode_i420(fr, mode != "yuv")
            if mode == "ppm":
                orc.rgb(out, fr.width, fr.height)
            if whole:
                one_pf.free()
    one(0)  # warm the caches and the page tables
    t0 = time.perf_counter()
    px = 0
    for _ in range(reps):
        for i, fr in enumerate(frames):
            one(i)
            px += fr.width * fr.height
    return px, time.perf_counter() - t0, reps * len(frames)


def _enc_worker(args):
    """The encoder row: enc_vp8_encode_i16x16_uv_sad_inloop of the reference (oracle/_ref/libref_enc.so), or of
    the oracle port where that is absent, on the benchmark's two 1920x1080 pictures."""
    reps, kind, search = args
    from encfix import EncOracle, EncReference, picture
    impl = EncReference() if kind == "reference" else EncOracle()
    pics = [picture(900, 1920, 1080, 0), picture(901, 1920, 1080, 1)]
    t0 = time.perf_counter()
    for _ in range(reps):
        for y, u, v in pics:
            impl.run(y, u, v, 75, search)
    return reps * len(pics) * 1920 * 1080, time.perf_counter() - t0, reps * len(pics)


def run_encoder(procs=None, seconds=4.0, search=1):
    from encfix import EncReference
    kind = "reference" if EncReference.available() else "port"
    procs = procs or len(os.sched_getaffinity(0)) or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        cal = pool.map(_enc_worker, [(1, kind, search)] * procs)
        reps = max(1, int(seconds / max(max(t for _, t, _ in cal), 1e-3)))
        res = pool.map(_enc_worker, [(reps, kind, search)] * procs)
    slowest = max(r[1] for r in res)
    return {"value": sum(r[0] for r in res) / slowest / 1e6, "unit": "Mpixel/s", "cores": procs, "kind": kind,
            "flags": "gcc -O3 (x86-64 baseline)" if kind == "reference" else "gcc -O2 (oracle/Makefile)",
            "sample": f"{sum(r[2] for r in res)} pictures ({reps} passes over 2 distinct 1920x1080 inputs on each of {procs} processes), {slowest:.1f} s, "
                      + ("enc_vp8_encode_bpred_uv_sad_inloop" if search == "bpred" else "enc_vp8_encode_i16x16_uv_sad_inloop") + ", quality 75"}


def run(files, mode="yuvf", procs=None, seconds=12.0, whole=False):
    from vp8fix import Reference
    kind = "reference" if Reference.available() else "port"
    procs = procs or len(os.sched_getaffinity(0)) or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        # calibrate with one repetition on every core at once (contention included), then size the sample
        cal = pool.map(_worker, [(files, mode, 1, kind, whole)] * procs)
        per_rep = max(t for _, t, _ in cal)
        reps = max(1, int(seconds / max(per_rep, 1e-3)))
        res = pool.map(_worker, [(files, mode, reps, kind, whole)] * procs)
    px = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    frames = sum(r[2] for r in res)
    return {
        "value": px / slowest / 1e6, "unit": "Mpixel/s", "cores": procs, "kind": kind,
        "flags": ref_library()[1] if kind == "reference" else "gcc -O2 (oracle/Makefile)",
        "frames_per_s": frames / slowest,
        "sample": f"{frames} frames ({reps} passes over {len(files)} distinct inputs on each of {procs} processes), "
                  f"{slowest:.1f} s, mode -{mode}, " + ("whole decoder: .webp bytes in memory -> pixels in memory (m01..m07)" if whole
                                                         else "token decode excluded"),
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="*")
    ap.add_argument("--encoder", action="store_true", help="time the reference encoder's i16 in-loop front end instead of the decoder")
    ap.add_argument("--encoder-bpred", action="store_true", help="... its 4x4 sub-block SAD front end")
    ap.add_argument("--mode", default="yuvf", choices=["yuv", "yuvf", "ppm"])
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=12.0)
    ap.add_argument("--whole", action="store_true", help="time the whole decoder (.webp bytes -> pixels), not just m06+m07")
    a = ap.parse_args()
    if a.encoder or a.encoder_bpred:
        print(json.dumps(run_encoder(a.procs or None, a.seconds, "bpred" if a.encoder_bpred else 1)))
    else:
        print(json.dumps(run(a.files, a.mode, a.procs or None, a.seconds, a.whole)))
