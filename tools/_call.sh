set -u
mkdir -p gpurun_out
for m in l2 smem mix; do
  ARESDB_B200_DENSE_ACC=$m python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$m.err | tee gpurun_out/bench_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
for m in smem l2; do
ARESDB_B200_DENSE_ACC=$m python bench.py --workload cfg3_count --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg3_count_$m.err | tee gpurun_out/bench_cfg3_count_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('count $m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
for m in l2 smem mix; do
ARESDB_B200_DENSE_ACC=$m python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg2_$m.err | tee gpurun_out/bench_cfg2_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 $m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
