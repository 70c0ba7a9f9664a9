#!/usr/bin/env python
"""Where a kernel's executed instructions and stall samples go, by SOURCE REGION of our own files.

ncu's source page attributes every SASS instruction to the innermost inlined line (a one-line helper in vp8_common.cuh, a
CUDA intrinsic header), which says little about a 2500-instruction macroblock step. This tool joins
  * the per-SASS-instruction counters of an .ncu-rep (ncu --page source --print-source sass), and
  * the inline call chains of the same kernel from `nvdisasm -gi` on the cubin inside libvp8gpu.so,
and bills each instruction to the outermost frame that lies in one of our kernel sources, then to the nearest preceding
`// ----` / `// ====` section comment of that file.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep vp8_mb_lockstep 'ILi4ELb1ELb1E' [--units 8355840]

`--units`: divide counts by this (e.g. macroblocks processed by the launch) to print per-unit figures.
The library must be the build the profile was taken from.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "webp-decoder_b200" / "csrc"
OURS = ("vp8_pairs_step_a.inc", "vp8_pairs_step_b.inc", "vp8_pairs_step_c.inc", "vp8_pairs_step_c1.inc", "vp8_pairs_step_c2.inc", "vp8_pairs_step_c3.inc", "vp8_pairs_row.inc", "vp8_pairs_image.inc", "vp8_pairs.cu")


def ncu_sass(rep: str, kernel: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    wi, xi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
    stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    insts = []
    for r in rows[hdr_i + 1:]:
        if len(r) <= ci or not r[0].startswith("0x"):
            continue
        insts.append({"sass": r[1].strip(), "n": int(r[ci] or 0), "smp": int(r[si] or 0), "wf": int(r[wi] or 0), "wfx": int(r[xi] or 0),
                      "stalls": {h: int(r[i] or 0) for h, i in stall_cols.items()}})
    return insts


def disasm_chains(lib: Path, func_key: str):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=td, capture_output=True, check=True)
        chains = None
        for cubin in Path(td).glob("*.cubin"):
            txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout
            lines = txt.split("\n")
            start = next((i for i, l in enumerate(lines) if l.startswith(".text.") and func_key in l), None)
            if start is None:
                continue
            chains, cur = [], None
            for l in lines[start + 1:]:
                if l.startswith("\t.section") or l.startswith(".text."):
                    break
                if "//## File" in l:
                    cur = [(Path(f).name, int(n)) for f, n in re.findall(r'"([^"]+)", line (\d+)', l)]
                elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
                    chains.append(cur)
            break
    if chains is None:
        sys.exit(f"no function matching {func_key!r} in {lib}")
    return chains


_src_cache = {}


def region(chain):
    if not chain:
        return "(no line info)"
    pick = None
    for f, n in chain:  # innermost first; keep the outermost frame in one of our kernel files
        if f in OURS:
            pick = (f, n)
    if pick is None:
        return chain[-1][0]
    f, n = pick
    if f not in _src_cache:
        _src_cache[f] = (CSRC / f).read_text().split("\n")
    src = _src_cache[f]
    for i in range(min(n, len(src)) - 1, -1, -1):
        if re.search(r"// (----|====)", src[i]):
            return f"{f}:{i + 1} {src[i].strip()[3:].strip(' -=')[:74]}"
    return f"{f}:top"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="regex for ncu -k")
    ap.add_argument("func_key", help="substring of the mangled function name that selects the template instance")
    ap.add_argument("--lib", default=str(ROOT / "webp-decoder_b200" / "libvp8gpu.so"))
    ap.add_argument("--units", type=float, default=0)
    args = ap.parse_args()
    insts = ncu_sass(args.report, args.kernel)
    chains = disasm_chains(Path(args.lib), args.func_key)
    if len(insts) != len(chains):
        sys.exit(f"instruction count mismatch: ncu {len(insts)} vs nvdisasm {len(chains)} (different build?)")
    tot_n = sum(i["n"] for i in insts)
    tot_s = sum(i["smp"] for i in insts)
    reg = collections.OrderedDict()
    for ins, ch in zip(insts, chains):
        r = reg.setdefault(region(ch), {"n": 0, "smp": 0, "static": 0, "wf": 0, "wfx": 0, "stalls": collections.Counter()})
        r["wf"] += ins["wf"]
        r["wfx"] += ins["wfx"]
        r["n"] += ins["n"]
        r["smp"] += ins["smp"]
        r["static"] += 1
        r["stalls"].update(ins["stalls"])
    print(f"{len(insts)} SASS instructions, {tot_n / 1e9:.3f} G executed (warp level), {tot_s} stall samples")
    unit = f"{'per unit':>9s}" if args.units else ""
    tot_wf = sum(i["wf"] for i in insts)
    tot_wfx = sum(i["wfx"] for i in insts)
    print(f"shared-memory wavefronts {tot_wf / 1e9:.3f} G, of which bank-conflict replays {tot_wfx / 1e9:.3f} G")
    print(f"{'static':>6s} {'executed':>9s} {unit} {'samples':>8s} {'smem wf':>8s} {'replays':>8s}  top stall reasons                          region")
    for name, r in reg.items():
        if r["n"] / max(tot_n, 1) < 0.002 and r["smp"] / max(tot_s, 1) < 0.002:
            continue
        top = ", ".join(f"{k[6:]} {v / max(r['smp'], 1) * 100:.0f}%" for k, v in r["stalls"].most_common(3))
        per = f"{r['n'] / args.units:9.1f}" if args.units else ""
        print(f"{r['static']:6d} {r['n'] / tot_n * 100:8.1f}% {per} {r['smp'] / max(tot_s, 1) * 100:7.1f}% {r['wf'] / max(tot_wf, 1) * 100:7.1f}% "
              f"{r['wfx'] / max(r['wf'], 1) * 100:7.1f}%  {top:42s} {name}")


if __name__ == "__main__":
    main()
