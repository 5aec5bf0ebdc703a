python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg3 cfg2 cfg4; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/bench_$w.err > gpurun_out/bench_$w.json; tail -1 gpurun_out/bench_$w.err; done
