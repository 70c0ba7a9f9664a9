#!/bin/bash
# usage: tools/ab_bench.sh <outdir> <batch> <tag>...   kernel-stage timing + all-frame parity of library variants (VP8_GPU_LIB)
out=$1; batch=$2; shift 2
mkdir -p $out
for t in "$@"; do
  lib=webp-decoder_b200/libvp8gpu_$t.so
  [ "$t" = default ] && lib=webp-decoder_b200/libvp8gpu.so
  VP8_GPU_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --batch $batch > $out/ab_${t}_$batch.json 2> $out/ab_${t}_$batch.err
  python - <<PY
import json
try:
    d = json.load(open("$out/ab_${t}_$batch.json"))
    print("$t", "$batch", round(d["ms_per_step"], 3), "ms", d["config"]["bit_exact_all_frames_vs_reference_digests"], d["config"]["launch"])
except Exception as e:
    print("$t", "failed", e)
PY
done
