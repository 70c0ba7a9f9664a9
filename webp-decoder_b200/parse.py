"""Host front end (include/vp8_parse.h): .webp bytes -> decoded frames, one image per host thread.

Counterpart of the reference's m01 + m02 + m05 call sequence in main.c:556-600
(webp_parse_simple_lossy -> vp8_parse_keyframe_header -> vp8_decode_decoded_frame).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import DecodedFrame, KeyFrameHeader

_L = None


def bind(L):
    global _L
    vp, sz = C.c_void_p, C.c_size_t
    L.vp8_parse_arena_bytes.argtypes = [C.c_uint32, C.c_uint32]
    L.vp8_parse_arena_bytes.restype = sz
    L.vp8_parse_webp_size.argtypes = [vp, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.vp8_parse_webp.argtypes = [vp, sz, vp, vp, vp, sz]
    L.vp8_parse_vp8.argtypes = [vp, sz, vp, vp, vp, sz]
    L.vp8_parse_free.argtypes = [vp]
    L.vp8_parse_free.restype = None
    L.vp8_parse_batch.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    _L = L


def _lib():
    if _L is None:
        from . import load_library
        load_library()
    return _L


class ParsedFrames:
    """n parsed frames: `.kfs[i]` / `.frames[i]` are the ctypes structs the GPU entry points take.
    The coefficient/mode arrays live either in calloc memory or in one pinned buffer (pinned=True)."""

    def __init__(self, n):
        self.n = n
        self.kfs = (KeyFrameHeader * n)()
        self.frames = (DecodedFrame * n)()
        self._pinned = None
        self._owned = True

    def free(self):
        if self._owned and self.frames is not None:
            L = _lib()
            for i in range(self.n):
                L.vp8_parse_free(C.addressof(self.frames[i]))
            self._owned = False
        if self._pinned is not None:
            self._pinned.close()
            self._pinned = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def kf_list(self):
        return [self.kfs[i] for i in range(self.n)]

    def frame_list(self):
        return [self.frames[i] for i in range(self.n)]

    def array(self, i, name):
        """numpy view of one per-frame array (for tests)."""
        f = self.frames[i]
        mb = f.mb_total
        count = {"segment_id": mb, "skip_coeff": mb, "has_coeff": mb, "ymode": mb, "uv_mode": mb, "bmode": mb * 16,
                 "coeff_y2": mb * 16, "coeff_y": mb * 256, "coeff_u": mb * 64, "coeff_v": mb * 64}[name]
        return np.ctypeslib.as_array(getattr(f, name), shape=(count,))


def webp_size(data: bytes):
    w, h = C.c_uint32(), C.c_uint32()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    if _lib().vp8_parse_webp_size(buf, len(data), C.byref(w), C.byref(h)) != 0:
        e = C.get_errno()
        raise OSError(e, "not a simple lossy WebP key frame")
    return w.value, h.value


def parse_batch(files, threads: int | None = None, pinned: bool = False) -> ParsedFrames:
    """Parse a list of .webp byte strings on `threads` host threads (default: all cores)."""
    L = _lib()
    n = len(files)
    assert n > 0
    threads = threads or os.cpu_count() or 1
    bufs = [(C.c_uint8 * len(d)).from_buffer_copy(d) for d in files]
    fp = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(d) for d in files])
    out = ParsedFrames(n)
    status = (C.c_int * n)()
    arenas = arena_sizes = None
    if pinned:
        from . import PinnedBuffer
        need = []
        for d in files:
            w, h = webp_size(d)
            need.append(int(L.vp8_parse_arena_bytes(w, h)))
        offs = np.concatenate([[0], np.cumsum(need)])
        out._pinned = PinnedBuffer(int(offs[-1]))
        base = out._pinned.array.ctypes.data
        arenas = (C.c_void_p * n)(*[base + int(o) for o in offs[:-1]])
        arena_sizes = (C.c_size_t * n)(*need)
    failed = L.vp8_parse_batch(fp, sizes, n, threads, out.kfs, out.frames, arenas, arena_sizes, status)
    if failed:
        bad = [(i, status[i]) for i in range(n) if status[i]]
        out.free()
        raise OSError(bad[0][1], f"{failed} of {n} files failed to parse; first: file {bad[0][0]}: {os.strerror(bad[0][1])}")
    return out


def parse_webp(data: bytes) -> ParsedFrames:
    return parse_batch([data], threads=1)
