"""Shared pieces of the encoder-row tests: seeded synthetic pictures, the CPU oracle
(oracle/vp8_enc_oracle.c) and, in the CPU container, the reference encoder library (oracle/_ref/libref_enc.so)."""
import ctypes as C
import hashlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"


def picture(seed, w, h, kind):
    """kind 0: uniform noise, 1: gradients with wrap-around edges, 2: gaussian texture on flat chroma, 3: flat."""
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if kind == 0:
        y, u, v = (rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), (ch, cw), (ch, cw)))
    elif kind == 1:
        yy, xx = np.mgrid[0:h, 0:w]
        y = ((xx * 3 + yy * 2 + seed) % 256).astype(np.uint8)
        yy, xx = np.mgrid[0:ch, 0:cw]
        u, v = ((xx * 5) % 256).astype(np.uint8), ((yy * 7 + 40) % 256).astype(np.uint8)
    elif kind == 2:
        y = np.clip(rng.normal(128, 30, (h, w)), 0, 255).astype(np.uint8)
        u, v = np.full((ch, cw), 90, np.uint8), np.clip(rng.normal(160, 8, (ch, cw)), 0, 255).astype(np.uint8)
    else:
        y, u, v = np.full((h, w), 17 + seed % 200, np.uint8), np.full((ch, cw), 128, np.uint8), np.full((ch, cw), 250, np.uint8)
    return tuple(np.ascontiguousarray(p) for p in (y, u, v))


def cases(count, seed0=0, max_w=150, max_h=120):
    """Seeded (seed, w, h, kind, quality) tuples: odd sizes, single rows / columns, every quality band."""
    rng = np.random.default_rng(1000 + seed0)
    out = [(seed0, 1, 1, 0, 50), (seed0 + 1, 16, 16, 0, 75), (seed0 + 2, 17, 33, 1, 0), (seed0 + 3, 150, 1, 2, 100), (seed0 + 4, 1, 90, 0, 10)]
    while len(out) < count:
        out.append((seed0 + len(out), int(rng.integers(1, max_w + 1)), int(rng.integers(1, max_h + 1)), int(rng.integers(0, 4)),
                    int(rng.integers(0, 101))))
    return out[:count]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class EncOracle:
    """oracle/liboracle.so: orc_enc_i16_inloop (our C restatement)."""

    def __init__(self):
        import subprocess
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "oracle"], check=True)
        self.lib = C.CDLL(str(so))
        self.lib.orc_enc_i16_inloop.argtypes = [C.c_void_p] * 3 + [C.c_uint32] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 6
        self.lib.orc_enc_bpred_inloop.argtypes = [C.c_void_p] * 3 + [C.c_uint32] * 4 + [C.c_int] + [C.c_void_p] * 7

    def run(self, y, u, v, quality, search, want_recon=False):
        """search: 0 DC, 1 whole-macroblock modes, "bpred" 4x4 sub-block modes."""
        h, w = y.shape
        mb = ((w + 15) // 16) * ((h + 15) // 16)
        co, ym, cm = np.zeros(mb * 400, np.int16), np.zeros(mb, np.uint8), np.zeros(mb, np.uint8)
        rec = [np.zeros(mb * k, np.uint8) for k in (256, 64, 64)] if want_recon else [None] * 3
        if search == "bpred":
            bm = np.zeros(mb * 16, np.uint8)
            qi = self.lib.orc_enc_bpred_inloop(y.ctypes.data, u.ctypes.data, v.ctypes.data, w, h, y.strides[0], u.strides[0], quality,
                                               ym.ctypes.data, bm.ctypes.data, cm.ctypes.data, co.ctypes.data,
                                               *[r.ctypes.data if r is not None else None for r in rec])
            assert qi >= 0
            out = {"coeffs": co, "y_modes": ym, "uv_modes": cm, "b_modes": bm, "qindex": qi}
            if want_recon:
                out.update(rec_y=rec[0], rec_u=rec[1], rec_v=rec[2])
            return out
        qi = self.lib.orc_enc_i16_inloop(y.ctypes.data, u.ctypes.data, v.ctypes.data, w, h, y.strides[0], u.strides[0], quality, int(search),
                                         ym.ctypes.data, cm.ctypes.data, co.ctypes.data, *[r.ctypes.data if r is not None else None for r in rec])
        assert qi >= 0
        out = {"coeffs": co, "y_modes": ym, "uv_modes": cm, "qindex": qi}
        if want_recon:
            out.update(rec_y=rec[0], rec_u=rec[1], rec_v=rec[2])
        return out


class EncReference:
    """The unmodified reference encoder modules (oracle/_ref/libref_enc.so; CPU container only)."""
    SO = ORACLE_DIR / "_ref" / "libref_enc.so"

    @classmethod
    def available(cls):
        return cls.SO.exists()

    class Img(C.Structure):  # reference src/enc-m04_yuv/enc_rgb_to_yuv.h:9-17
        _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("y_stride", C.c_uint32), ("uv_stride", C.c_uint32),
                    ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p)]

    def __init__(self):
        self.lib = C.CDLL(str(self.SO))
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def run(self, y, u, v, quality, search):
        h, w = y.shape
        img = self.Img(w, h, y.strides[0], u.strides[0], y.ctypes.data, u.ctypes.data, v.ctypes.data)
        co, n, qi = C.POINTER(C.c_int16)(), C.c_size_t(), C.c_uint8()
        if search == "bpred":
            P8 = C.POINTER(C.c_uint8)
            pym, nym, pbm, nbm, pcm, ncm = P8(), C.c_size_t(), P8(), C.c_size_t(), P8(), C.c_size_t()
            rc = self.lib.enc_vp8_encode_bpred_uv_sad_inloop(C.byref(img), quality, C.byref(pym), C.byref(nym), C.byref(pbm), C.byref(nbm),
                                                             C.byref(pcm), C.byref(ncm), C.byref(co), C.byref(n), C.byref(qi))
            assert rc == 0
            out = {"coeffs": np.ctypeslib.as_array(co, (n.value,)).copy(), "y_modes": np.ctypeslib.as_array(pym, (nym.value,)).copy(),
                   "uv_modes": np.ctypeslib.as_array(pcm, (ncm.value,)).copy(), "b_modes": np.ctypeslib.as_array(pbm, (nbm.value,)).copy(),
                   "qindex": qi.value}
            for ptr in (pym, pbm, pcm, co):
                self.libc.free(ptr)
            return out
        if not search:
            rc = self.lib.enc_vp8_encode_dc_pred_inloop(C.byref(img), quality, C.byref(co), C.byref(n), C.byref(qi))
            assert rc == 0
            mb = n.value // 400
            ym, cm = np.zeros(mb, np.uint8), np.zeros(mb, np.uint8)
        else:
            pym, nym, pcm, ncm = C.POINTER(C.c_uint8)(), C.c_size_t(), C.POINTER(C.c_uint8)(), C.c_size_t()
            rc = self.lib.enc_vp8_encode_i16x16_uv_sad_inloop(C.byref(img), quality, C.byref(pym), C.byref(nym), C.byref(pcm), C.byref(ncm),
                                                              C.byref(co), C.byref(n), C.byref(qi))
            assert rc == 0
            ym, cm = np.ctypeslib.as_array(pym, (nym.value,)).copy(), np.ctypeslib.as_array(pcm, (ncm.value,)).copy()
            self.libc.free(pym), self.libc.free(pcm)
        c = np.ctypeslib.as_array(co, (n.value,)).copy()
        self.libc.free(co)
        return {"coeffs": c, "y_modes": ym, "uv_modes": cm, "qindex": qi.value}


def same(a, b, search):
    if search == "bpred" and not np.array_equal(a["b_modes"], b["b_modes"]):
        return False
    return (a["qindex"] == b["qindex"] and np.array_equal(a["coeffs"], b["coeffs"])
            and (not search or (np.array_equal(a["y_modes"], b["y_modes"]) and np.array_equal(a["uv_modes"], b["uv_modes"]))))
