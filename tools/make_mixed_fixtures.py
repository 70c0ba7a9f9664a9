#!/usr/bin/env python
"""BASELINE config 5 inputs, made OFFLINE with the reference's own tools (oracle/_ref/gen_ppm, encoder, decoder; needs
/root/reference at build time, see oracle/Makefile) and committed: bench_data/mixed/*.webp + digests.json (sha256 of the
reference decoder's -yuv / -yuvf / -ppm bytes). Sizes 129x129 .. 3840x2160 incl. non-multiples of 16, q 10 / 50 / 95,
encoder modes i16 and bpred, all gen_ppm patterns (noise only where the file stays small).

    python tools/make_mixed_fixtures.py            (about two minutes on 8 cores)
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
OUT = ROOT / "bench_data" / "mixed"
sys.path.insert(0, str(ROOT / "tools"))
from config5_sweep import ppm_to_png  # noqa: E402

SIZES = [(129, 129), (256, 256), (512, 512), (1000, 700), (1280, 720), (1920, 1080), (2560, 1440), (3840, 2160)]
QS = [10, 50, 95]
MODES = ["i16", "bpred"]
PATTERNS = ["noise", "rgbgrad", "checker", "diag"]


def jobs():
    k = 0
    for (w, h) in SIZES:
        for q in QS:
            for mode in MODES:
                # noise is incompressible: keep it to the sizes / qualities where the file stays small
                pats = [p for p in PATTERNS if p != "noise" or w * h <= 512 * 512 or (q == 10 and w <= 1280)]
                yield pats[k % len(pats)], w, h, q, mode, 100 + k
                k += 1


def make(job):
    pattern, w, h, q, mode, seed = job
    stem = f"{pattern}_{w}x{h}_q{q}_{mode}"
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        ppm, png, webp = td / "a.ppm", td / "a.png", OUT / (stem + ".webp")
        subprocess.run([str(REF / "gen_ppm"), pattern, str(w), str(h), str(ppm), str(seed)], check=True, capture_output=True)
        ppm_to_png(ppm, png)
        subprocess.run([str(REF / "encoder"), "--q", str(q), "--loopfilter", "--mode", mode, str(png), str(webp)], check=True, capture_output=True)
        d = {"width": w, "height": h, "q": q, "mode": mode, "pattern": pattern, "bytes": webp.stat().st_size}
        for flag, key in (("-yuv", "yuv"), ("-yuvf", "yuvf"), ("-ppm", "ppm")):
            o = td / ("o." + key)
            subprocess.run([str(REF / "decoder"), flag, str(webp), str(o)], check=True, capture_output=True)
            d[key] = hashlib.sha256(o.read_bytes()).hexdigest()
    return stem + ".webp", d


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(8) as ex:
        res = dict(ex.map(make, list(jobs())))
    (OUT / "digests.json").write_text(json.dumps(res, indent=1, sort_keys=True))
    print(len(res), "files,", sum(v["bytes"] for v in res.values()), "bytes")
