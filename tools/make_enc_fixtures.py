#!/usr/bin/env python
"""Writes tests/golden/enc.json: sha256 of what the UNMODIFIED reference encoder modules (oracle/_ref/libref_enc.so, built from
/root/reference by oracle/Makefile) return for seeded synthetic pictures - enc_vp8_encode_dc_pred_inloop and
enc_vp8_encode_i16x16_uv_sad_inloop / enc_vp8_encode_bpred_uv_sad_inloop (coefficients, modes, qindex). The GPU box has no /root/reference: there the tests
compare the CUDA path and the CPU oracle with these digests. Run in the CPU container: python tools/make_enc_fixtures.py"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from encfix import GOLDEN, EncReference, cases, digest, picture  # noqa: E402

ref = EncReference()
out = {}
big = [(900, 1920, 1080, 0, 75), (901, 1920, 1080, 1, 75), (902, 1280, 720, 2, 30), (903, 1000, 700, 0, 90)]
for seed, w, h, kind, q in cases(48) + big:
    y, u, v = picture(seed, w, h, kind)
    for search in (0, 1, "bpred"):
        r = ref.run(y, u, v, q, search)
        arrays = [r["coeffs"], r["y_modes"], r["uv_modes"]] + ([r["b_modes"]] if search == "bpred" else [])
        out[f"{seed}_{w}x{h}_k{kind}_q{q}_s{2 if search == 'bpred' else search}"] = {"digest": digest(*arrays), "qindex": r["qindex"]}
(GOLDEN / "enc.json").write_text(json.dumps(out, indent=0, sort_keys=True))
print(len(out), "digests ->", GOLDEN / "enc.json")
