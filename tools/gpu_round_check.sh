#!/bin/bash
# One GPU session that produces everything the round's evidence needs (run under gpurun from the repo root):
# the GPU test-suite, smoke(), the default bench line (device-resident + e2e + cpu baseline), the reference arm, the
# other BASELINE configs, the ncu launch list of the bench command and full captures of the dominant kernels.
# FULL=1 adds the configs / captures whose kernels did not change in the last iterations.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_reference.err | tee gpurun_out/bench_reference.json | cut -c1-200
WL="cfg2 cfg3_count cfg4"
[ "${FULL:-0}" = 1 ] && WL="$WL cfg4_hll"
for w in $WL; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | cut -c1-200
done
[ "${FULL:-0}" = 1 ] && python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-zone-maps 2>gpurun_out/bench_nozm.err | tee gpurun_out/bench_nozm.json | cut -c1-200
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --profile-range > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/prof_fused_jit \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/prof_fused_jit_cfg4 \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_cfg4.log 2>&1
if [ "${FULL:-0}" = 1 ]; then
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/prof_fused_jit_count \
    python bench.py --workload cfg3_count --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_count.log 2>&1
fi
tail -2 gpurun_out/ncu_full.log
