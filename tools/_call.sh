set -u
mkdir -p gpurun_out
for m in mix 3; do
  ARESDB_B200_DENSE_ACC=$m python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
python bench.py --workload cfg3_count --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg3_count.err | tee gpurun_out/bench_cfg3_count.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('count', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
for m in mix 3; do
ARESDB_B200_DENSE_ACC=$m python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg2_$m.err | tee gpurun_out/bench_cfg2_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 $m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/prof_fused_dense_mix \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/prof_fused_dense_count \
    python bench.py --workload cfg3_count --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_count.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --profile-range > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/ncu_full.log
