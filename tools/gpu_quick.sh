#!/bin/bash
# parity tests + a short probe (no profiling)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/probe_step.py > gpurun_out/probe.log 2>&1
grep '"group": 0' gpurun_out/probe.log
