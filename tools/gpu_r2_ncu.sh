#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each), round 2.  gpurun brings back at most 64 MiB:
# every capture is exported to CSV on the box (raw metrics + source page) and only two .ncu-rep files are kept.
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --verify-queries 2"
cap() { # name workload kernel-regex skip
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2_$1 python bench.py --workload $2 $B > gpurun_out/r2_ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i gpurun_out/r2_$1.ncu-rep --page raw --csv > gpurun_out/r2_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_$1.ncu-rep --page source --csv > gpurun_out/r2_$1_source.csv 2>/dev/null
}
cap coarse_c2 c2 coarse_pair_kernel 2
cap finalize_c2 c2 finalize_kernel 2
cap coarse_c3 c3 coarse_pair_kernel 5
cap coarse_c4 c4 coarse_pair_kernel 5
rm -f gpurun_out/r2_finalize_c2.ncu-rep gpurun_out/r2_coarse_c4.ncu-rep
du -sh gpurun_out
