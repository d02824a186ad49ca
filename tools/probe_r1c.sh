#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/quick_perf_widths.py > gpurun_out/quick_perf_widths.txt 2>&1; cat gpurun_out/quick_perf_widths.txt
