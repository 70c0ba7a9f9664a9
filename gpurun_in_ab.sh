for k in 3 2; do
  timeout 300 python tools/gpu_check.py --quick --kernel $k > gpurun_out/check_k$k.log 2>&1; r1=$?
  timeout 200 python bench.py --kernel $k --no-cpu-baseline --no-e2e > gpurun_out/bench_k$k.json 2> gpurun_out/bench_k$k.err; r2=$?
  echo "kernel $k check=$r1 $(grep -cE 'ALL OK' gpurun_out/check_k$k.log) fails=$(grep -c FAIL gpurun_out/check_k$k.log) bench=$r2 $(python -c "import json;d=json.load(open('gpurun_out/bench_k$k.json'));print(d['ms_per_step'], d['config']['launch']['smem_bytes'], d['config']['parity_spot_check_vs_reference_digests'])")"
done
