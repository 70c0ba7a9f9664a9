"""The CPU oracle (oracle/vp8_oracle.c) against everything that pins it:
   - digests produced by the UNMODIFIED reference decoder binary on 254 .webp inputs (tests/golden/digests.json),
   - RGB pixels of the reference's dwebp-derived golden PNGs (149 of those inputs), i.e. libwebp itself,
   - digests of the reference hot path on 120 struct-level fuzz frames (simple filter, sharpness, lf deltas, int16 wrap),
   - and, where oracle/_ref exists (the CPU container), the reference library called directly on fresh fuzz frames."""
import ctypes as C
import json

import numpy as np
import pytest

from vp8fix import GOLDEN, Frame, fuzz_frame, sha


class ParsedAsFrame:
    """Adapter: a (kf, DecodedFrame) pair from the product parser with the Frame interface Oracle expects."""

    def __init__(self, kf, d):
        self.width, self.height, self._d = kf.width, kf.height, d
        self.mb_cols, self.mb_rows = d.mb_cols, d.mb_rows

    def cstruct(self, drop_has_coeff=False):
        from vp8fix import DecodedFrame
        c = DecodedFrame.from_buffer_copy(bytes(self._d))
        c._keep = self._d
        return c

    @property
    def i420_size(self):
        return self.width * self.height + 2 * ((self.width + 1) // 2) * ((self.height + 1) // 2)


def test_oracle_matches_reference_decoder_digests(oracle, golden, parsed_golden):
    assert len(golden) >= 250
    for name, g in golden.items():
        kf, d, _ = parsed_golden[name]
        fr = ParsedAsFrame(kf, d)
        w, h = g["width"], g["height"]
        assert (fr.width, fr.height) == (w, h), name
        yuv, yuvf = oracle.decode_i420(fr, False), oracle.decode_i420(fr, True)
        assert sha(yuv) == g["yuv"], f"{name}: -yuv"
        assert sha(yuvf) == g["yuvf"], f"{name}: -yuvf"
        rgb = oracle.rgb(yuvf, w, h)
        assert sha(oracle.ppm(rgb, w, h)) == g["ppm"], f"{name}: -ppm"
        assert sha(oracle.png(rgb, w, h)) == g["png"], f"{name}: -png"
        if "dwebp_rgb" in g:
            assert sha(rgb) == g["dwebp_rgb"], f"{name}: libwebp golden PNG"


def test_golden_corpus_covers_the_branches(golden):
    infos = [g["info"] for g in golden.values()]
    assert sum(i["Use segment"] == "1" for i in infos) >= 40
    assert sum(i["Level"] == "0" for i in infos) >= 5
    assert sum(int(i["MB B_PRED"]) > 0 for i in infos) >= 60
    assert sum(0 < int(i["MB B_PRED"]) < int(i["MB total"]) for i in infos) >= 20  # mixed i16 / B_PRED frames
    assert sum("dwebp_rgb" in g for g in golden.values()) == 149


def test_oracle_matches_reference_fuzz_digests(oracle):
    fz = json.loads((GOLDEN / "fuzz.json").read_text())
    assert len(fz) == 120
    seen = dict(simple=0, sharp=0, delta=0, seg=0)
    for seed, g in fz.items():
        fr = fuzz_frame(int(seed), g["width"], g["height"], **g["kw"])
        seen["simple"] += fr.params["lf_use_simple"]
        seen["sharp"] += fr.params["lf_sharpness"] != 0
        seen["delta"] += fr.params["lf_delta_enabled"]
        seen["seg"] += fr.params["segmentation_enabled"]
        yuvf = oracle.decode_i420(fr, True)
        assert sha(oracle.decode_i420(fr, False)) == g["yuv"], seed
        assert sha(yuvf) == g["yuvf"], seed
        assert sha(oracle.ppm(oracle.rgb(yuvf, fr.width, fr.height), fr.width, fr.height)) == g["ppm"], seed
    assert min(seen.values()) >= 30, seen  # the branches no bitstream fixture reaches are exercised


@pytest.mark.parametrize("seed", range(0, 60))
def test_oracle_vs_reference_library(oracle, reference, seed):
    rng = np.random.default_rng(seed + 777)
    w, h = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    fr = fuzz_frame(seed + 5000, w, h, amp=[5, 40, 400, 2500][seed % 4], density=[0.05, 0.2, 0.6][seed % 3], raw=bool(seed % 2))
    for filt in (False, True):
        assert np.array_equal(oracle.decode_i420(fr, filt), reference.decode_i420(fr, filt))
    # has_coeff == NULL is legal input (vp8_loopfilter.c:226)
    assert np.array_equal(oracle.decode_i420(fr, True, drop_has_coeff=True), reference.decode_i420(fr, True, drop_has_coeff=True))
    yuvf = oracle.decode_i420(fr, True)
    rgb = oracle.rgb(yuvf, w, h)
    assert oracle.ppm(rgb, w, h) == reference.ppm_bytes(yuvf, w, h)
    assert oracle.png(rgb, w, h) == reference.png_bytes(yuvf, w, h)
    # stand-alone loop filter on padded planes
    y, u, v = oracle.recon_padded(fr)
    y2, u2, v2 = y.copy(), u.copy(), v.copy()
    oracle.loopfilter_padded(fr, y, u, v)
    reference.loopfilter_padded(fr, y2, u2, v2)
    assert np.array_equal(y, y2) and np.array_equal(u, u2) and np.array_equal(v, v2)


def test_rgb_closed_form_on_tiny_and_odd_sizes(oracle, reference):
    rng = np.random.default_rng(5)
    for w, h in [(1, 1), (2, 2), (3, 5), (16, 16), (17, 9), (18, 10), (31, 32), (64, 7), (2, 1), (1, 2), (5, 1)]:
        i420 = rng.integers(0, 256, w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)).astype(np.uint8)
        assert np.array_equal(oracle.rgb(i420, w, h), reference.rgb(i420, w, h)), (w, h)


def test_png_block_boundaries(oracle, reference):
    # raw scanline stream sizes around the 65535-byte stored-block limit
    rng = np.random.default_rng(9)
    for w, h in [(21845, 1), (150, 145), (149, 146), (2731, 8), (300, 73)]:
        i420 = rng.integers(0, 256, w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)).astype(np.uint8)
        assert oracle.png(oracle.rgb(i420, w, h), w, h) == reference.png_bytes(i420, w, h), (w, h)
