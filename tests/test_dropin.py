"""The drop-in itself: the reference's UNMODIFIED main.c + m01..m05, linked against libvp8gpu.so instead of its
m06..m09 objects (oracle/Makefile target _ref/decoder_gpu). Its -yuv / -yuvf / -ppm / -png files must be
byte-identical to the reference decoder's (digests in tests/golden/digests.json)."""
import hashlib
import subprocess

import pytest

from vp8fix import GOLDEN, REF_DIR

pytestmark = pytest.mark.gpu


def test_reference_cli_linked_against_libvp8gpu(golden, tmp_path):
    exe = REF_DIR / "decoder_gpu"
    if not exe.exists():
        pytest.skip("oracle/_ref/decoder_gpu not built (needs /root/reference at build time)")
    names = [n for n in sorted(golden) if golden[n]["width"] * golden[n]["height"] >= 64 * 40]
    picks = names[::40][:3] + ["enc_noise_1x1_q10_rdo_lf.webp" if "enc_noise_1x1_q10_rdo_lf.webp" in golden else names[0]]
    checked = 0
    for n in picks:
        for flag, key in (("-yuv", "yuv"), ("-yuvf", "yuvf"), ("-ppm", "ppm"), ("-png", "png")):
            out = tmp_path / f"out.{key}"
            r = subprocess.run([str(exe), flag, str(GOLDEN / "webp" / n), str(out)], capture_output=True, text=True)
            assert r.returncode == 0, (n, flag, r.stderr)
            assert hashlib.sha256(out.read_bytes()).hexdigest() == golden[n][key], (n, flag)
            checked += 1
    assert checked == 4 * len(picks) >= 12
