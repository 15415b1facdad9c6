#!/bin/bash
# parity tests + ncu full capture of the step kernel (3000 triangles, 1080p) + probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 5 -c 2 -f -o gpurun_out/prof_eval \
    python tools/profile_step.py --steps 10 > gpurun_out/prof_eval.log 2>&1
python tools/probe_step.py > gpurun_out/probe.log 2>&1
grep '"group": 0' gpurun_out/probe.log
