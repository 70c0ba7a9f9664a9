"""Python face of include/vp8_enc.h: the reference encoder's in-loop reconstruction for whole-macroblock prediction, on the GPU.

    encode_dc_pred_inloop(y, u, v, quality)          reference src/enc-m08_recon/enc_recon.h:69-73
    encode_i16x16_uv_sad_inloop(y, u, v, quality)    reference src/enc-m08_recon/enc_recon.h:97-105
    encode_bpred_uv_sad_inloop(y, u, v, quality)     reference src/enc-m08_recon/enc_recon.h:120-130
    encode_batch(pictures, quality, search)          many pictures per launch (vp8_gpu_enc_i16_inloop)

Planes are 2-D uint8 numpy arrays (Y: h x w, U and V: ceil(h/2) x ceil(w/2)). Plumbing only: every number comes from the
CUDA kernel; without the library or a device the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import Vp8GpuError, load_library


class EncYuv420Image(C.Structure):
    """reference src/enc-m04_yuv/enc_rgb_to_yuv.h:9-17"""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("y_stride", C.c_uint32), ("uv_stride", C.c_uint32),
                ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p)]


def _lib():
    L = load_library()
    if not getattr(L, "_enc_ready", False):
        vp, pp = C.c_void_p, C.POINTER(C.c_void_p)
        L.vp8_gpu_enc_i16_inloop.argtypes = [C.c_int, pp, C.c_int, C.c_int, C.c_int, pp, pp, pp, pp, pp, pp, C.POINTER(C.c_uint8)]
        L.vp8_gpu_enc_bpred_inloop.argtypes = [C.c_int, pp, C.c_int, C.c_int, pp, pp, pp, pp, pp, pp, pp, C.POINTER(C.c_uint8)]
        L.enc_vp8_encode_bpred_uv_sad_inloop.argtypes = [vp, C.c_int, pp, C.POINTER(C.c_size_t), pp, C.POINTER(C.c_size_t), pp,
                                                         C.POINTER(C.c_size_t), pp, C.POINTER(C.c_size_t), C.POINTER(C.c_uint8)]
        L.vp8_gpu_enc_mb_total.argtypes = [C.c_uint32, C.c_uint32]
        L.vp8_gpu_enc_mb_total.restype = C.c_size_t
        L.vp8_gpu_enc_last_kernel_ms.restype = C.c_double
        L.enc_vp8_encode_dc_pred_inloop.argtypes = [vp, C.c_int, pp, C.POINTER(C.c_size_t), C.POINTER(C.c_uint8)]
        L.enc_vp8_encode_i16x16_uv_sad_inloop.argtypes = [vp, C.c_int, pp, C.POINTER(C.c_size_t), pp, C.POINTER(C.c_size_t), pp,
                                                          C.POINTER(C.c_size_t), C.POINTER(C.c_uint8)]
        L._enc_ready = True
    return L


def _image(y, u, v):
    for p in (y, u, v):
        if p.dtype != np.uint8 or p.ndim != 2 or p.strides[1] != 1:
            raise ValueError("planes must be 2-D uint8 arrays with contiguous rows")
    h, w = y.shape
    if u.shape != ((h + 1) // 2, (w + 1) // 2) or v.shape != u.shape or u.strides[0] != v.strides[0]:
        raise ValueError("chroma planes must be ceil(h/2) x ceil(w/2) with one stride")
    return EncYuv420Image(w, h, y.strides[0], u.strides[0], y.ctypes.data, u.ctypes.data, v.ctypes.data)


def _take(ptr, count, dtype):
    """Copies a callee-malloc'ed array and frees it, as a caller of the reference does."""
    n = count * np.dtype(dtype).itemsize
    out = np.frombuffer(C.string_at(ptr, n), dtype=dtype).copy()
    C.CDLL(None).free(C.c_void_p(ptr))
    return out


def _check(rc, what):
    if rc != 0:
        e = C.get_errno()
        raise Vp8GpuError(e, f"{what}: {load_library().vp8_gpu_last_error().decode()}")


def encode_dc_pred_inloop(y, u, v, quality):
    """-> (coeffs int16[mb_total * 400], qindex)."""
    L = _lib()
    img = _image(y, u, v)
    co, n, qi = C.c_void_p(), C.c_size_t(), C.c_uint8()
    _check(L.enc_vp8_encode_dc_pred_inloop(C.byref(img), quality, C.byref(co), C.byref(n), C.byref(qi)), "enc_vp8_encode_dc_pred_inloop")
    return _take(co.value, n.value, np.int16), qi.value


def encode_i16x16_uv_sad_inloop(y, u, v, quality):
    """-> (y_modes uint8[mb_total], uv_modes uint8[mb_total], coeffs int16[mb_total * 400], qindex)."""
    L = _lib()
    img = _image(y, u, v)
    ym, nym, cm, ncm, co, n, qi = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_uint8()
    _check(L.enc_vp8_encode_i16x16_uv_sad_inloop(C.byref(img), quality, C.byref(ym), C.byref(nym), C.byref(cm), C.byref(ncm), C.byref(co),
                                                 C.byref(n), C.byref(qi)), "enc_vp8_encode_i16x16_uv_sad_inloop")
    return _take(ym.value, nym.value, np.uint8), _take(cm.value, ncm.value, np.uint8), _take(co.value, n.value, np.int16), qi.value


def encode_bpred_uv_sad_inloop(y, u, v, quality):
    """-> (y_modes, b_modes uint8[mb_total * 16], uv_modes, coeffs, qindex)."""
    L = _lib()
    img = _image(y, u, v)
    ym, nym, bm, nbm, cm, ncm, co, n, qi = (C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(),
                                             C.c_size_t(), C.c_uint8())
    _check(L.enc_vp8_encode_bpred_uv_sad_inloop(C.byref(img), quality, C.byref(ym), C.byref(nym), C.byref(bm), C.byref(nbm), C.byref(cm),
                                                C.byref(ncm), C.byref(co), C.byref(n), C.byref(qi)), "enc_vp8_encode_bpred_uv_sad_inloop")
    return (_take(ym.value, nym.value, np.uint8), _take(bm.value, nbm.value, np.uint8), _take(cm.value, ncm.value, np.uint8),
            _take(co.value, n.value, np.int16), qi.value)


def encode_batch(pictures, quality, search, device=-1, want_recon=False, out_buffer=None):
    """pictures: [(y, u, v)]. search: 0 DC, 1 whole-macroblock modes, "bpred" 4x4 sub-block modes.
    -> list of dicts {coeffs, y_modes, uv_modes[, b_modes][, rec_y, rec_u, rec_v]} and the qindex.
    out_buffer: optional uint8 array (e.g. a PinnedBuffer's, batch_out_bytes() long) the results are laid out in; copies into
    pinned memory are asynchronous."""
    L = _lib()
    n = len(pictures)
    imgs = [_image(*p) for p in pictures]
    arr = (C.c_void_p * n)(*[C.addressof(i) for i in imgs])
    bpred = search == "bpred"
    outs, cols = [], {k: (C.c_void_p * n)() for k in ("coeffs", "y_modes", "uv_modes", "b_modes", "rec_y", "rec_u", "rec_v")}
    at = 0

    def take(nbytes, dtype):
        nonlocal at
        if out_buffer is None:
            return np.empty(nbytes // np.dtype(dtype).itemsize, dtype)
        a = out_buffer[at:at + nbytes].view(dtype)
        at += (nbytes + 255) // 256 * 256
        return a
    for i, (y, _, _) in enumerate(pictures):
        h, w = y.shape
        mb = L.vp8_gpu_enc_mb_total(w, h)
        o = {"coeffs": take(mb * 800, np.int16), "y_modes": take(mb, np.uint8), "uv_modes": take(mb, np.uint8)}
        if bpred:
            o["b_modes"] = take(mb * 16, np.uint8)
        if want_recon:
            o.update(rec_y=take(mb * 256, np.uint8), rec_u=take(mb * 64, np.uint8), rec_v=take(mb * 64, np.uint8))
        for k, a in o.items():
            cols[k][i] = a.ctypes.data
        outs.append(o)
    qi = C.c_uint8()
    rec = [cols[k] if want_recon else None for k in ("rec_y", "rec_u", "rec_v")]
    if bpred:
        _check(L.vp8_gpu_enc_bpred_inloop(device, arr, n, quality, cols["coeffs"], cols["y_modes"], cols["b_modes"], cols["uv_modes"], *rec,
                                          C.byref(qi)), "vp8_gpu_enc_bpred_inloop")
    else:
        _check(L.vp8_gpu_enc_i16_inloop(device, arr, n, quality, 1 if search else 0, cols["coeffs"], cols["y_modes"], cols["uv_modes"], *rec,
                                        C.byref(qi)), "vp8_gpu_enc_i16_inloop")
    return outs, qi.value


def batch_out_bytes(pictures, want_recon=False, bpred=False):
    """Bytes encode_batch needs in out_buffer."""
    L = _lib()
    total = 0
    for y, _, _ in pictures:
        mb = L.vp8_gpu_enc_mb_total(y.shape[1], y.shape[0])
        for nbytes in (mb * 800, mb, mb) + ((mb * 16,) if bpred else ()) + ((mb * 256, mb * 64, mb * 64) if want_recon else ()):
            total += (nbytes + 255) // 256 * 256
    return total


def last_kernel_ms():
    return _lib().vp8_gpu_enc_last_kernel_ms()
