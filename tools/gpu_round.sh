#!/bin/bash
# One GPU visit: parity tests, bench (both arms), ncu launch list + full capture of the step kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc > gpurun_out/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 5 -c 3 -f -o gpurun_out/prof_eval \
    python tools/profile_step.py --steps 10 > gpurun_out/prof_eval.log 2>&1
python tools/probe_step.py > gpurun_out/probe.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
