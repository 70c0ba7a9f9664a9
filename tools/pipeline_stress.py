#!/usr/bin/env python
"""Repeats the body of tests/test_gpu_parity.py::test_pipelined_decode_matches_digests (all golden inputs through the
pipelined calls in chunks of 7: three transports x -yuvf / -yuv / -ppm x pinned / pageable output) for a number of seconds and
reports the first wrong frame in detail (which call, which frame, how many bytes differ and where).
    python tools/pipeline_stress.py [seconds]"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import webp_decoder_b200 as W  # noqa: E402
from webp_decoder_b200 import parse as P  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
golden = json.loads((ROOT / "tests" / "golden" / "digests.json").read_text())
names = sorted(golden)
pf = P.parse_batch([(ROOT / "tests" / "golden" / "webp" / n).read_bytes() for n in names])
kfs, frs = [pf.kfs[i] for i in range(pf.n)], [pf.frames[i] for i in range(pf.n)]
ctx = W.Context(0)
good = {}
t0, rounds, calls = time.time(), 0, 0
while time.time() - t0 < seconds:
    for compact in (True, False, "auto"):
        ctx.set_transport(compact, 3)
        for ppm, key, filtered in ((False, "yuvf", True), (False, "yuv", False), (True, "ppm", True)):
            need = ctx.decode_bytes(kfs, ppm=ppm)
            pinned = W.PinnedBuffer(need)
            for kind, out in (("pinned", pinned.array), ("pageable", np.empty(need, np.uint8))):
                out[:] = 0xAA
                offs, sizes = ctx.decode_into(kfs, frs, out, filtered=filtered, ppm=ppm, chunk=7)
                calls += 1
                for i, (n, o, s) in enumerate(zip(names, offs, sizes)):
                    got = out[int(o):int(o) + int(s)]
                    if hashlib.sha256(got).hexdigest() != golden[n][key]:
                        ref = good.get((key, n))
                        where = "no good copy kept"
                        if ref is not None and len(ref) == len(got):
                            d = np.flatnonzero(ref != got)
                            where = f"{len(d)} of {len(got)} bytes differ, first at {d[:8].tolist()}, last at {int(d[-1])}; got {got[d[:8]].tolist()} want {ref[d[:8]].tolist()}"
                        print(f"MISMATCH round {rounds} call {calls}: transport {compact} {key} {kind} output, frame {i} {n} "
                              f"{golden[n]['width']}x{golden[n]['height']}: {where}", flush=True)
                        sys.exit(1)
                    elif rounds == 0 and kind == "pinned":
                        good[(key, n)] = got.copy()
            pinned.close()
    rounds += 1
print(f"pipeline stress ok: {rounds} rounds, {calls} pipelined calls, {calls * len(names)} frames")
