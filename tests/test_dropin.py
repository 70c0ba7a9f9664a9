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


def test_reference_encoder_cli_linked_against_libvp8gpu(oracle, tmp_path):
    """The encoder-side drop-in (oracle/Makefile target _ref/encoder_gpu): the reference's unmodified encoder_main.c and
    modules with the DC, whole-macroblock and sub-block SAD in-loop front ends taken from libvp8gpu.so. Its .webp files for
    --mode dc, --mode i16 and --mode bpred must equal the reference encoder's byte for byte; --mode bpred-rdo (still the
    reference's own code in that binary) shows the rest of the program is untouched."""
    import numpy as np
    gpu, ref = REF_DIR / "encoder_gpu", REF_DIR / "encoder"
    if not gpu.exists() or not ref.exists():
        pytest.skip("oracle/_ref/encoder_gpu not built (needs /root/reference at build time)")
    rng = np.random.default_rng(11)
    checked = 0
    for (w, h), q in (((77, 45), 60), ((16, 16), 5), ((130, 97), 90), ((1, 1), 50)):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q != 90 else np.tile(np.arange(w, dtype=np.uint8)[None, :, None] * 2, (h, 1, 3))
        src = tmp_path / f"in_{w}x{h}.png"
        src.write_bytes(bytes(oracle.png(np.ascontiguousarray(rgb).reshape(-1), w, h)))
        for mode in ("dc", "i16", "bpred", "bpred-rdo"):
            outs = []
            for exe in (ref, gpu):
                out = tmp_path / f"{exe.name}_{mode}.webp"
                r = subprocess.run([str(exe), "--q", str(q), "--mode", mode, "--loopfilter", str(src), str(out)], capture_output=True, text=True)
                assert r.returncode == 0, (exe.name, mode, r.stderr)
                outs.append(out.read_bytes())
            assert outs[0] == outs[1], (w, h, q, mode)
            checked += 1
    assert checked == 16
