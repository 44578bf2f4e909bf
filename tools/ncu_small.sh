#!/bin/bash
# HBM-bound kernels under ncu: packed scan and streaming exact scan at one query (run on a B200 box)
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum.per_second,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size"
ncu --metrics $M --clock-control none -k regex:scan_packed16 -c 2 --csv --log-file gpurun_out/r1_packed_q1.csv python tools/quick_packed.py 50000000 1 32 hamming > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:scan_packed16 -c 1 --csv --log-file gpurun_out/r1_packed_q256.csv python tools/quick_packed.py 50000000 256 32 hamming > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:scan_stream -c 2 --csv --log-file gpurun_out/r1_stream_q1.csv python bench.py --nq 1 --steps 1 --warmup 1 --plan exact --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:coarse_single -c 2 --csv --log-file gpurun_out/r1_single_q64.csv python bench.py --nq 64 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
