#!/bin/bash
# round-2 session f: all GPU tests (new: preprocessing, PixelShuffle heads, C4 at size, plugin parts 5-7)
set -u
mkdir -p gpurun_out
timeout -k 5 1800 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_f.log
cat gpurun_out/plugin_gpu_report.txt 2>/dev/null | tail -2
