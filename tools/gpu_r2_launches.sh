#!/bin/bash
mkdir -p gpurun_out
for spec in "c2:--rows 1250000:c2shard" "c2::c2" "c3::c3" "c4::c4"; do
  w=${spec%%:*}; rest=${spec#*:}; extra=${rest%%:*}; tag=${rest#*:}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$tag.csv python bench.py --workload $w $extra --steps 2 --warmup 1 --no-cpu-baseline --verify-queries 4 > gpurun_out/r2_launches_$tag.log 2>&1
  echo "== $tag rc=$?"
done
