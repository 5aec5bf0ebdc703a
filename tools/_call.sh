set -u
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@"; }
run --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n2 fixed', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'], d['e2e'])"
ARESDB_B200_EXCHANGE=exact run --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_n2_exact.err | tee gpurun_out/bench_n2_exact.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n2 exact', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'])"
run --workload cfg4 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_n2_cfg4.err | tee gpurun_out/bench_n2_cfg4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n2 cfg4', d['ms_per_step'], d['roofline']['kernel_ms'], d['groups'], d['gpu_launches'])"
tail -3 gpurun_out/bench_n2.err
