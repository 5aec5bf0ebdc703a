set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for m in mix smem; do
ARESDB_B200_DENSE_ACC=$m python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_sum_$m.err | tee gpurun_out/bench_sum_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sum $m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'])"
done
python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg4.err | tee gpurun_out/bench_cfg4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'])"
for m in mix smem; do
ARESDB_B200_DENSE_ACC=$m python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg2_$m.err | tee gpurun_out/bench_cfg2_$m.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 $m', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
done
