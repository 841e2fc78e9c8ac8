#!/bin/bash
# One gpurun call that produces the raw material of profiles/ for a round: each program first exits 0 without ncu,
# then runs once under ncu (one GPU).  Outputs go to gpurun_out/; scripts/ncu_summary.py and
# scripts/launch_list_summary.py turn them into the committed summaries.
# usage (from the repo root, on the GPU box): bash scripts/capture_profiles.sh r02
R=${1:-r02}
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
set -x
MODEL=m71 python scripts/gpu_one_mh.py > $O/${R}_one_m71_plain.log 2>&1 && \
  MODEL=m71 $NCU -k regex:'mh_kernel|loglik_kernel' --launch-skip 2 --launch-count 2 -o $O/${R}_mh_m71 python scripts/gpu_one_mh.py > $O/${R}_one_m71_ncu.log 2>&1
MODEL=gauss python scripts/gpu_one_mh.py > $O/${R}_one_gauss_plain.log 2>&1 && \
  MODEL=gauss $NCU -k regex:'mh_kernel|loglik_kernel' --launch-skip 2 --launch-count 2 -o $O/${R}_mh_gauss python scripts/gpu_one_mh.py > $O/${R}_one_gauss_ncu.log 2>&1
CARRY=1 python scripts/gpu_one_mh.py > $O/${R}_one_carry_plain.log 2>&1 && \
  CARRY=1 $NCU -k regex:'mh_kernel' --launch-skip 1 --launch-count 1 -o $O/${R}_mh_carry python scripts/gpu_one_mh.py > $O/${R}_one_carry_ncu.log 2>&1
python scripts/gpu_aggregate.py > $O/${R}_agg_plain.log 2>&1 && \
  $NCU -k regex:agg_mh_kernel --launch-skip 4 --launch-count 4 -o $O/${R}_agg python scripts/gpu_aggregate.py > $O/${R}_agg_ncu.log 2>&1
python bench.py --tiles-per-gpu 148 --field-tiles 148 --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_bench148.json 2> $O/${R}_bench148.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${R}_launches.csv \
    python bench.py --tiles-per-gpu 148 --field-tiles 148 --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_launches_ncu.log 2>&1
# the merge back is limited to 64 MiB: summarise on the box, keep only the main kernel's report
for k in mh_m71 mh_gauss mh_carry agg; do
  [ -f $O/${R}_$k.ncu-rep ] && python scripts/ncu_summary.py $O/${R}_$k.ncu-rep $O/${R}_ncu_$k.txt "$R $k" \
    && ncu -i $O/${R}_$k.ncu-rep --page raw --csv | gzip > $O/${R}_${k}_raw.csv.gz
done
rm -f $O/${R}_mh_gauss.ncu-rep $O/${R}_mh_carry.ncu-rep $O/${R}_agg.ncu-rep
ls -la $O | grep ${R}_
