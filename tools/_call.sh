set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_sum.err | tee gpurun_out/bench_sum.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sum', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'])"
python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
python bench.py --workload cfg3_count --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_cfg3_count.err | tee gpurun_out/bench_cfg3_count.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('count', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'])"
