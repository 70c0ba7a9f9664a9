#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of the box, one direction at a time and both at once.
The end-to-end number of bench.py is bounded by the device->host leg (3.11 MB per 1080p frame); this probe gives the
floor it is compared with in DESIGN.md. Run under gpurun: python tools/pcie_probe.py"""
import json
import torch

def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    res["h2d_GBps"] = n / timed(lambda: d_in.copy_(h_in, non_blocking=True)) / 1e6
    res["d2h_GBps"] = n / timed(lambda: h_out.copy_(d_out, non_blocking=True)) / 1e6
    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    ms = timed(both)
    res["both_directions_each_GBps"] = n / ms / 1e6
    # 64 chunks of 16 MiB, the way a pipelined call issues them
    def chunks():
        for i in range(0, n, 1 << 24):
            h_out[i:i + (1 << 24)].copy_(d_out[i:i + (1 << 24)], non_blocking=True)
    res["d2h_16MiB_chunks_GBps"] = n / timed(chunks) / 1e6
    print(json.dumps(res))

if __name__ == "__main__":
    main()
