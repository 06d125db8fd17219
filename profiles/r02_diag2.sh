#!/bin/bash
mkdir -p gpurun_out/d4
LGR_LOOP_DEBUG=1 LGR_REFERENCE_ROOT=scratch/reference timeout 600 python profiles/log_loop_gpu.py --iters 2 > gpurun_out/d4/log.txt 2>&1
grep "debug\]" gpurun_out/d4/log.txt | cut -c1-900
