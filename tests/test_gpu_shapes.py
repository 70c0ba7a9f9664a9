"""Parity at the shapes that are timed and reported: the benchmark batch itself, the BASELINE
config-5 matrix, and a seeded soak of random batch / launch shapes. All through the C-ABI, bytes compared with the
reference decoder's digests (bench_data/) or with the CPU oracle."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

from vp8fix import sha

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _frames_equal_digests(buf, offs, sizes, order, want):
    """Frame k of a replicated batch must hash to want[order[k]]: the first copy of each distinct input is hashed, every
    further copy is compared with it byte by byte (same statement, far less hashing)."""
    first = {}
    bad = []
    for k, i in enumerate(order):
        got = buf[int(offs[k]):int(offs[k]) + int(sizes[k])]
        if i not in first:
            first[i] = got
            if hashlib.sha256(got).hexdigest() != want[i]:
                bad.append(k)
        elif not np.array_equal(got, first[i]):
            bad.append(k)
    return bad


def test_benchmark_batch_every_frame_yuv_yuvf_ppm(lib, gpu_ctx):
    """BASELINE config 2 / 4 exactly as bench.py times it: 1024 x 1080p (the four bench_data inputs x 256, interleaved) in one
    launch of the lockstep kernel, 7 images per CTA - all 1024 frames of -yuv, -yuvf and -ppm against the reference
    decoder's digests."""
    from webp_decoder_b200 import parse as P
    names = ["noise_1920x1080_q75.webp", "rgbgrad_1920x1080_q75.webp", "checker_1920x1080_q75.webp", "diag_1920x1080_q75.webp"]
    dg = json.loads((ROOT / "bench_data" / "digests.json").read_text())
    pf = P.parse_batch([(ROOT / "bench_data" / n).read_bytes() for n in names], pinned=True)
    order = [k % 4 for k in range(1024)]
    b = gpu_ctx.upload([pf.kfs[i] for i in order], [pf.frames[i] for i in order])
    try:
        for filtered, key in ((False, "yuv"), (True, "yuvf")):
            gpu_ctx.run(b, filtered, lib.TIGHT)
            cfg = gpu_ctx.last_launch_config()
            assert cfg["images_per_cta"] == 7 and cfg["warps_per_image"] == 4 and cfg["segments"] == 1, cfg  # vp8_mb_lockstep, one wave
            buf, offs, sizes = gpu_ctx.download_i420(b)
            assert not _frames_equal_digests(buf, offs, sizes, order, [dg[n][key] for n in names]), key
            del buf
        gpu_ctx.rgb(b)
        buf, offs, sizes = gpu_ctx.download_ppm(b)
        assert not _frames_equal_digests(buf, offs, sizes, order, [dg[n]["ppm"] for n in names])
    finally:
        b.free()
        pf.free()
        gpu_ctx.trim()


def test_config5_matrix_sizes_qualities_modes(lib, gpu_ctx):
    """BASELINE config 5: 129x129 .. 3840x2160 (non-multiples of 16 included) x q 10 / 50 / 95 x encoder modes i16 / bpred x
    every gen_ppm pattern (bench_data/mixed, made offline with the reference's encoder by tools/make_mixed_fixtures.py):
    one mixed-size batch through the kernels, and through the three pipelined calls, against the reference decoder's
    -yuv / -yuvf / -ppm digests."""
    from webp_decoder_b200 import parse as P
    mdir = ROOT / "bench_data" / "mixed"
    dg = json.loads((mdir / "digests.json").read_text())
    names = sorted(dg)
    assert len(names) == 48 and {(dg[n]["q"], dg[n]["mode"]) for n in names} == {(q, m) for q in (10, 50, 95) for m in ("i16", "bpred")}
    datas = [(mdir / n).read_bytes() for n in names]
    pf = P.parse_batch(datas, pinned=True)
    kfs, frs = pf.kf_list(), pf.frame_list()
    for filtered, key in ((False, "yuv"), (True, "yuvf")):
        outs = gpu_ctx.decode_i420(kfs, frs, filtered=filtered)
        assert not [n for n, o in zip(names, outs) if sha(o) != dg[n][key]], key
    assert not [n for n, p in zip(names, gpu_ctx.decode_ppm(kfs, frs)) if sha(p) != dg[n]["ppm"]]
    need = gpu_ctx.decode_bytes(kfs)
    out = lib.PinnedBuffer(need)
    # dense contract (balanced transport), compact frames, .webp bytes
    cf = P.parse_batch_compact(datas, pinned=True)
    wf = lib.WebpFiles(datas)
    for label, call in (("dense", lambda: gpu_ctx.decode_into(kfs, frs, out.array, filtered=True)),
                        ("compact", lambda: gpu_ctx.decode_compact_into(cf.frame_list(), out.array, filtered=True)),
                        ("webp", lambda: gpu_ctx.decode_webp_into(wf, out.array, filtered=True))):
        out.array[:] = 0
        offs, sizes = call()
        assert not [n for n, o, s in zip(names, offs, sizes) if sha(out.array[int(o):int(o) + int(s)]) != dg[n]["yuvf"]], label
    out.close()
    cf.free()
    pf.free()
    gpu_ctx.trim()


def test_seeded_soak_of_batch_and_launch_shapes(lib, gpu_ctx, oracle):
    """tools/soak.py with a fixed seed: 150 rounds of random batch sizes, frame sizes, content density, warps per image,
    images per SM, cluster sizes, both schedules and every entry point, each frame compared with the oracle. The
    wavefront's progress stamps are plain shared / global memory traffic ordered by fences (no sanitizer on this pool),
    so this is the race check that runs with the test set."""
    sys.path.insert(0, str(ROOT / "tools"))
    import soak
    try:
        rounds, frames = soak.soak(gpu_ctx, oracle, np.random.default_rng(20251018), rounds=150)
    finally:
        gpu_ctx.set_kernel(3)
        gpu_ctx.set_tuning(0, 0)
        gpu_ctx.set_cluster(0, True)
        gpu_ctx.set_transport("auto", 0)
    assert rounds == 150 and frames > 3000
