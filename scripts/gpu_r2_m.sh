#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python scripts/hot_trace.py 64 4096 > gpurun_out/r2m_trace64.txt 2>&1; cat gpurun_out/r2m_trace64.txt | tail -28
timeout 300 python scripts/hot_trace.py 128 8192 > gpurun_out/r2m_trace128.txt 2>&1; cat gpurun_out/r2m_trace128.txt | tail -28
