#!/bin/bash
# Quick GPU visit while iterating on the step kernels: parity tests, lane-group sweep, bench, per-kernel times.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/probe_step.py > gpurun_out/probe.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
TPOSE_CUDA_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/kernels.csv python tools/profile_step.py --steps 20 > gpurun_out/kernels.log 2>&1
