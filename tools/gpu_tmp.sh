#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/prof_cfg5_entropy.py 64 grad > gpurun_out/r02_cfg5_entropy_grad_w64.txt 2>&1; tail -16 gpurun_out/r02_cfg5_entropy_grad_w64.txt
timeout 120 python tools/prof_cfg5_entropy.py 64 > gpurun_out/r02_cfg5_entropy_w64.txt 2>&1; head -3 gpurun_out/r02_cfg5_entropy_w64.txt | grep -v Warn
