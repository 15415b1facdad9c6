#!/bin/bash
# One GPU visit: parity tests, bench (both arms), ncu launch list + full capture of the step kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc > gpurun_out/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# every launch of a short bench run with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --batch 4 > gpurun_out/bench_under_ncu.log 2>&1
# the dominant kernel, full set, single image
TPOSE_CUDA_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_edges -s 5 -c 3 -f -o gpurun_out/prof_edges \
    python tools/profile_step.py --steps 10 > gpurun_out/prof_edges.log 2>&1
TPOSE_CUDA_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_corners -s 5 -c 2 -f -o gpurun_out/prof_corners \
    python tools/profile_step.py --steps 10 > gpurun_out/prof_corners.log 2>&1
# DRAM traffic of the batched step (64 images per launch)
TPOSE_CUDA_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:"k_edges|k_corners" -s 8 -c 4 --csv --log-file gpurun_out/batched_traffic.csv \
    python tools/probe_batch.py --batches 64 --steps 3 > gpurun_out/batched_traffic.log 2>&1
python tools/probe_step.py > gpurun_out/probe.log 2>&1
python tools/probe_batch.py --batches 1,2,3,4,8,32,64,128 > gpurun_out/probe_batch.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python tests/run_configs.py --quick > gpurun_out/configs_quick.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
