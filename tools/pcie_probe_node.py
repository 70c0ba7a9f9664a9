#!/usr/bin/env python
"""Aggregate pinned host<->device bandwidth of the NODE: the probe of tools/pcie_probe.py on every GPU at the same time
(one process per GPU, both directions busy for about two seconds). The end-to-end numbers of `bench.py --gpus N` cannot
exceed (bytes each rank moves) / (this node figure / N). Run under `gpurun --gpus N`: python tools/pcie_probe_node.py N"""
import json
import subprocess
import sys
import time

WORKER = r"""
import sys, time, torch
n = 1 << 29
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
start = float(sys.argv[1])
while time.time() < start: pass
t0 = time.time(); reps = 0
while time.time() - t0 < 2.0:
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); reps += 1
dt = time.time() - t0
print(reps * n / dt / 1e9)
"""


def main():
    import os
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    res = {}
    for group in ([0], list(range(n))) if n > 1 else ([0],):
        start = time.time() + 25  # every process has imported torch and pinned its buffers by then
        procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(start)], stdout=subprocess.PIPE, text=True,
                                  env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(g))) for g in group]
        each = [float(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        res[f"{len(group)}_gpus_busy"] = {"each_direction_GBps_per_gpu": [round(x, 1) for x in each],
                                          "node_total_GBps_both_directions": round(2 * sum(each), 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
