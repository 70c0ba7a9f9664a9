#!/bin/bash
# usage: tools/ncu_traffic.sh <outdir> <tag>...   DRAM bytes + duration of one vp8_mb_lockstep launch on the benchmark batch, per library variant
out=$1; shift
mkdir -p $out
for t in "$@"; do
  lib=webp-decoder_b200/libvp8gpu_$t.so
  [ "$t" = default ] && lib=webp-decoder_b200/libvp8gpu.so
  export VP8_GPU_LIB=$PWD/$lib
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > $out/plain_$t.json 2> $out/plain_$t.err &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:vp8_mb_lockstep -s 2 -c 1 --csv \
      --log-file $out/traffic_$t.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > $out/ncu_$t.log 2>&1
  echo "== $t"; grep -E "dram__bytes|gpu__time|inst_executed" $out/traffic_$t.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
