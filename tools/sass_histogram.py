#!/usr/bin/env python
"""SASS opcode histogram of one kernel from an .ncu-rep: static count, executed share, executions per unit.
    python tools/sass_histogram.py gpurun_out/prof.ncu-rep vp8_mb_lockstep 8355840 > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, kernel, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kernel}"], capture_output=True,
                     text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = rows[h].index("Instructions Executed")
st, ex, full = collections.Counter(), collections.Counter(), collections.Counter()
name = next((r[1] for r in rows[:h] if r and r[0] == "Kernel Name"), kernel)
for r in rows[h + 1:]:
    if len(r) <= ci or not r[0].startswith("0x"):
        continue
    toks = r[1].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    st[base] += 1
    ex[base] += int(r[ci] or 0)
    if base in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDGSTS", "REDUX", "SHFL", "BAR", "ATOMS", "MEMBAR", "VOTE", "WARPSYNC", "UTMALDG", "UBLKCP"):
        full[op] += 1
tot = sum(ex.values())
print(f"{name} (sm_100a) - SASS opcode histogram from {rep}")
print(f"{sum(st.values())} instructions static, {tot / 1e9:.3f} G executed = {tot / units:.0f} per unit\n")
print(f"{'opcode':12s} {'static':>7s} {'executed':>9s} {'per unit':>9s}")
for op, n in sorted(ex.items(), key=lambda kv: -kv[1]):
    print(f"{op:12s} {st[op]:7d} {n / tot * 100:8.1f}% {n / units:9.1f}")
print("\nmemory / special opcodes by full mnemonic (static):")
for op, n in sorted(full.items()):
    print(f"  {op:28s} {n}")
tma = [op for op in st if op.startswith(("UTMALDG", "UBLKCP", "UTC", "LDTM", "STTM"))]
print("\nTMA / tcgen05 opcodes: " + (", ".join(tma) if tma else "none (expected: not a contraction; coefficient staging is cp.async = LDGSTS, SURVEY.md 2.2)"))
