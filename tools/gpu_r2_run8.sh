#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/f16_scan_probe.py > gpurun_out/r2_f16_scan.txt 2>&1; cat gpurun_out/r2_f16_scan.txt
