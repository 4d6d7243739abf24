#!/bin/bash
# one gpurun call (end of round 2): launch list of the bench step + --set full captures of the hot kernels
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 2 --warmup 1 --only main --no-cpu-baseline > gpurun_out/r02f_launches_bench.log 2>&1
for spec in "build:corr_build_tc_staged" "build60:corr_build_tc_staged" "lookup:corr_lookup" "neus:neus_forward" \
            "mapping:neus_grid_bwd" "mapping:neus_composite_bwd"; do
  mode=${spec%%:*}; kern=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/r02f_${mode}_${kern} \
      python tools/profile_driver.py $mode 3 > gpurun_out/r02f_ncu_${mode}_${kern}.log 2>&1
done
ls -la gpurun_out/ | grep r02f
