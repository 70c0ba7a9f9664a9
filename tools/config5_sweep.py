#!/usr/bin/env python
"""BASELINE config 5 in small: sizes x qualities x encoder modes x patterns, encoded on the spot with the reference's own
tools (oracle/_ref/gen_ppm, oracle/_ref/encoder), decoded by the reference decoder binary (oracle/_ref/decoder -yuvf,
one process per core) and by the GPU path (one mixed-size batch through vp8_gpu_run, and once more through the pipelined
vp8_gpu_decode_i420), byte-compared. Prints one JSON object. Run under gpurun:  python tools/config5_sweep.py [--big]"""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W  # noqa: E402
from webp_decoder_b200 import parse as P  # noqa: E402


def ppm_to_png(ppm: Path, png: Path):
    data = ppm.read_bytes()
    parts, pos = [], 0
    while len(parts) < 4:  # P6, width, height, maxval
        end = pos
        while data[end:end + 1] not in (b" ", b"\n", b"\t", b"\r"):
            end += 1
        parts.append(data[pos:end])
        pos = end + 1
    w, h = int(parts[1]), int(parts[2])
    rgb = np.frombuffer(data, np.uint8, w * h * 3, pos).reshape(h, w * 3)
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb], axis=1).tobytes()  # filter type 0 per row

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    png.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                    chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))


def make_one(job):
    tmp, pattern, w, h, q, mode, seed = job
    stem = f"{pattern}_{w}x{h}_q{q}_{mode}"
    ppm, png, webp, ref = tmp / (stem + ".ppm"), tmp / (stem + ".png"), tmp / (stem + ".webp"), tmp / (stem + ".i420")
    subprocess.run([str(REF / "gen_ppm"), pattern, str(w), str(h), str(ppm), str(seed)], check=True, capture_output=True)
    ppm_to_png(ppm, png)
    subprocess.run([str(REF / "encoder"), "--q", str(q), "--loopfilter", "--mode", mode, str(png), str(webp)], check=True, capture_output=True)
    subprocess.run([str(REF / "decoder"), "-yuvf", str(webp), str(ref)], check=True, capture_output=True)
    digest = hashlib.sha256(ref.read_bytes()).hexdigest()
    for f in (ppm, png, ref):
        f.unlink()
    return stem, webp, digest, w, h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="add 1440p and 4K")
    args = ap.parse_args()
    sizes = [(256, 256), (129, 129), (1000, 700), (1280, 720), (1920, 1080)] + ([(2560, 1440), (3840, 2160)] if args.big else [])
    jobs = []
    with tempfile.TemporaryDirectory() as td:
        tmp = Path(td)
        for (w, h) in sizes:
            for q in (10, 50, 90):
                for mode in ("i16", "bpred"):
                    for pattern in ("noise", "rgbgrad"):
                        jobs.append((tmp, pattern, w, h, q, mode, 1 + len(jobs)))
        t0 = time.time()
        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            made = list(ex.map(make_one, jobs))
        t_make = time.time() - t0
        blobs = [m[1].read_bytes() for m in made]
        pf = P.parse_batch(blobs, pinned=True)
        ctx = W.Context(0)
        kfs, frs = pf.kf_list(), pf.frame_list()
        outs = ctx.decode_i420(kfs, frs, filtered=True)
        bad = [m[0] for m, o in zip(made, outs) if hashlib.sha256(bytes(o)).hexdigest() != m[2]]
        buf = np.empty(ctx.decode_bytes(kfs), np.uint8)
        offs, szs = ctx.decode_into(kfs, frs, buf, filtered=True, chunk=16)
        bad_p = [m[0] for m, o, s in zip(made, offs, szs) if hashlib.sha256(buf[int(o):int(o) + int(s)].tobytes()).hexdigest() != m[2]]
        b = ctx.upload(kfs, frs)
        for _ in range(3):
            ctx.run(b, True, W.TIGHT)
        ctx.kernel_time()
        for _ in range(10):
            ctx.run(b, True, W.TIGHT)
        ms, n = ctx.kernel_time()
        b.free()
        px = sum(m[3] * m[4] for m in made)
        print(json.dumps({"images": len(made), "sizes": sizes, "q": [10, 50, 90], "modes": ["i16", "bpred"], "patterns": ["noise", "rgbgrad"],
                          "encode_and_reference_decode_s": round(t_make, 1), "mismatches_batch": bad, "mismatches_pipelined": bad_p,
                          "kernel_ms_per_mixed_batch": ms / n, "mpixel_per_s_kernel_stage": px / (ms / n) / 1e3,
                          "launch": ctx.last_launch_config()}))
        pf.free()
        ctx.close()
        return 1 if bad or bad_p else 0


if __name__ == "__main__":
    sys.exit(main())
