#!/bin/bash
# One GPU session that produces the round's evidence (run under gpurun from the repo root): the GPU test-suite, smoke(),
# the default bench line (device-resident + e2e + cpu baseline + the other BASELINE configs as sub-results), the reference
# arm, the ncu launch list of the bench command and full captures of the dominant kernels.  R = file prefix.
set -u
R=${R:-r02_final}
mkdir -p gpurun_out
(python -m pytest tests -m gpu -q 2>&1 | tail -4) | tee gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${R}_smoke.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/${R}_bench.err | tee gpurun_out/${R}_bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${R}_bench_reference.err | tee gpurun_out/${R}_bench_reference.json | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub --profile-range > gpurun_out/${R}_launches_cfg3.log 2>&1
for w in cfg4 cfg4_hll; do
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub --profile-range > gpurun_out/${R}_launches_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/${R}_fused_dense_sum \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub > gpurun_out/${R}_ncu_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/${R}_fused_hll \
    python bench.py --workload cfg4_hll --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub > gpurun_out/${R}_ncu_hll.log 2>&1
if [ "${FULL:-0}" = 1 ]; then
ncu --set full --clock-control none --import-source on -k regex:aresFusedJit -s 10 -c 1 -f -o gpurun_out/${R}_fused_global_cfg4 \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub > gpurun_out/${R}_ncu_cfg4.log 2>&1
fi
tail -1 gpurun_out/${R}_ncu_hll.log | cut -c1-120
