#!/usr/bin/env python
"""Regenerates tests/golden/ and bench_data/ from the UNMODIFIED reference (run in the CPU container only).

Needs /root/reference and oracle/_ref (make -C oracle ref). Everything it writes is committed, so neither is
needed to RUN the tests or the benchmark:

  tests/golden/webp/*.webp      inputs: (a) gen_ppm pattern -> scripts/ppm_to_png.py -> reference encoder
                                (--q Q --loopfilter [--mode M]); (b) a sample of the reference's own fixture corpora
                                (cwebp-made streams: segmentation, i16/B_PRED mixes, libwebp probabilities)
  tests/golden/digests.json     per input: sha256 of the reference decoder's -yuv, -yuvf, -ppm and -png bytes, and
                                for the reference's dwebp-derived golden PNGs the sha256 of their RGB pixels
  tests/golden/fuzz.json        sha256 of the reference's -yuv/-yuvf/-ppm output on struct-level fuzz frames
                                (tests/vp8fix.fuzz_frame seeds): pins simple filter, sharpness, lf deltas, int16 wrap
  bench_data/*.webp             1080p / 512x512 / 4K benchmark inputs (BASELINE.json configs), digests in
                                bench_data/digests.json
"""
from __future__ import annotations

import hashlib
import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from vp8fix import GOLDEN, REF_DIR, Reference, fuzz_frame, sha  # noqa: E402

REF = Path("/root/reference")
TMP = Path(tempfile.mkdtemp(prefix="vp8fx"))


def run(*cmd):
    subprocess.run([str(c) for c in cmd], check=True)


def encode(pattern, w, h, seed, q, mode, lf, out: Path):
    ppm, png = TMP / "in.ppm", TMP / "in.png"
    run(REF_DIR / "gen_ppm", pattern, w, h, ppm, seed)
    run(sys.executable, REF / "scripts" / "ppm_to_png.py", ppm, png)
    cmd = [REF_DIR / "encoder", "--q", q]
    if lf:
        cmd.append("--loopfilter")
    if mode:
        cmd += ["--mode", mode]
    run(*cmd, png, out)


def ref_digests(webp: Path):
    d = {}
    for flag, key in (("-yuv", "yuv"), ("-yuvf", "yuvf"), ("-ppm", "ppm"), ("-png", "png")):
        out = TMP / ("out." + key)
        run(REF_DIR / "decoder", flag, webp, out)
        d[key] = hashlib.sha256(out.read_bytes()).hexdigest()
    info = subprocess.run([str(REF_DIR / "decoder"), "-info", str(webp)], capture_output=True, text=True, check=True).stdout
    kv = {ln.split(":")[0].strip(): ln.split(":")[1].strip() for ln in info.splitlines() if ":" in ln}
    d["width"], d["height"] = int(kv["Width"]), int(kv["Height"])
    d["info"] = {k: kv[k] for k in ("Use segment", "Simple filter", "Level", "Sharpness", "Use lf delta", "Base Q", "MB B_PRED", "MB total") if k in kv}
    return d


def main():
    assert REF.exists() and (REF_DIR / "decoder").exists(), "run `make -C oracle ref` in the CPU container first"
    wdir = GOLDEN / "webp"
    if wdir.exists():
        shutil.rmtree(wdir)
    wdir.mkdir(parents=True)
    digests = {}

    # (a) our own encodes, small sizes around macroblock edges
    sizes = [(16, 16), (17, 17), (31, 33), (48, 48), (64, 40), (129, 129), (200, 120), (1, 1), (15, 70)]
    combos = []
    for i, (w, h) in enumerate(sizes):
        for j, pattern in enumerate(["noise", "rgbgrad", "checker", "diag"]):
            if w < 2 and pattern == "rgbgrad":
                continue  # gen_ppm divides by (w-1)
            q = [10, 50, 75, 95][(i + j) % 4]
            mode = [None, "bpred", "i16", "dc"][(i + 2 * j) % 4]
            lf = (i + j) % 5 != 0
            combos.append((pattern, w, h, 1 + i + j, q, mode, lf))
    for pattern, w, h, seed, q, mode, lf in combos:
        name = f"enc_{pattern}_{w}x{h}_q{q}_{mode or 'rdo'}_{'lf' if lf else 'nolf'}.webp"
        encode(pattern, w, h, seed, q, mode, lf, wdir / name)
        digests[name] = ref_digests(wdir / name)
        digests[name]["origin"] = f"gen_ppm {pattern} {w} {h} seed {seed} | encoder --q {q} {'--loopfilter ' if lf else ''}{'--mode ' + mode if mode else ''}"

    # (b) the reference's own corpora: all of images/webp + images/testimages/webp (with dwebp golden PNGs), a sample of generated/
    from PIL import Image
    picks = []
    for sub, png_sub in (("webp", "png-out"), ("testimages/webp", "testimages/png")):
        for f in sorted((REF / "images" / sub).glob("*.webp")):
            picks.append((f, REF / "images" / png_sub / (f.stem + ".png")))
    gen = sorted((REF / "images" / "generated" / "webp").glob("*.webp"))
    picks += [(f, None) for f in gen[::4]]
    for f, png in picks:
        name = "ref_" + f.parent.parent.name + "_" + f.name
        shutil.copyfile(f, wdir / name)
        digests[name] = ref_digests(wdir / name)
        digests[name]["origin"] = "reference fixture images/" + str(f.relative_to(REF / "images"))
        if png is not None and png.exists():
            digests[name]["dwebp_rgb"] = sha(Image.open(png).convert("RGB").tobytes())
    (GOLDEN / "digests.json").write_text(json.dumps(digests, indent=0, sort_keys=True))
    print("golden inputs:", len(digests), "bytes:", sum(p.stat().st_size for p in wdir.iterdir()))

    # (c) struct-level fuzz digests from the real reference hot path
    ref = Reference()
    fz = {}
    for seed in range(120):
        rng = np.random.default_rng(10_000 + seed)
        w, h = int(rng.integers(1, 140)), int(rng.integers(1, 140))
        kw = dict(amp=[5, 40, 400, 2500][seed % 4], density=[0.05, 0.2, 0.6][seed % 3], raw=bool(seed % 2))
        fr = fuzz_frame(seed, w, h, **kw)
        yuv, yuvf = ref.decode_i420(fr, False), ref.decode_i420(fr, True)
        fz[str(seed)] = dict(width=w, height=h, kw=kw, yuv=sha(yuv), yuvf=sha(yuvf), ppm=sha(ref.ppm_bytes(yuvf, w, h)))
    (GOLDEN / "fuzz.json").write_text(json.dumps(fz, indent=0, sort_keys=True))

    # (d) benchmark inputs
    bdir = ROOT / "bench_data"
    bdir.mkdir(exist_ok=True)
    bd = {}
    bench = [
        ("noise", 1920, 1080, 3, 75, None), ("rgbgrad", 1920, 1080, 1, 75, None), ("checker", 1920, 1080, 1, 75, None),
        ("diag", 1920, 1080, 1, 75, None),
        ("noise", 512, 512, 7, 75, None), ("rgbgrad", 512, 512, 1, 75, None),
        ("checker", 3840, 2160, 1, 75, None), ("rgbgrad", 3840, 2160, 1, 75, None),
    ]
    for pattern, w, h, seed, q, mode in bench:
        name = f"{pattern}_{w}x{h}_q{q}.webp"
        encode(pattern, w, h, seed, q, mode, True, bdir / name)
        bd[name] = ref_digests(bdir / name)
        bd[name]["origin"] = f"gen_ppm {pattern} {w} {h} seed {seed} | encoder --q {q} --loopfilter"
        print(name, (bdir / name).stat().st_size)
    (bdir / "digests.json").write_text(json.dumps(bd, indent=0, sort_keys=True))
    shutil.rmtree(TMP)


if __name__ == "__main__":
    main()
