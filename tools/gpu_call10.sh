#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_FFTS=2048,4096 timeout 200 python tools/gpu_render_ab.py 600 > gpurun_out/r01_v25_render_ab.txt 2>&1
tail -6 gpurun_out/r01_v25_render_ab.txt
