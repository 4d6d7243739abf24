#!/bin/bash
# one gpurun call: launch list of the bench step + --set full captures of the hot kernels + write microbenchmark
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --only main --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
for spec in "build:corr_build_tc" "build60:corr_build_tc" "lookup:corr_lookup" "ba:ba_persistent" "neus:neus_forward" \
            "balarge:ba_solve_cluster" "balarge:ba_system" "balarge:ba_linearize"; do
  mode=${spec%%:*}; kern=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/r02_${mode}_${kern} \
      python tools/profile_driver.py $mode 3 > gpurun_out/r02_ncu_${mode}_${kern}.log 2>&1
done
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/wbench tools/wbench.cu && /tmp/wbench > gpurun_out/r02_wbench.txt 2>&1
compute-sanitizer --tool racecheck --racecheck-report all python -m pytest "tests/test_gpu_ba_large.py::test_ba_large_vs_oracle[P31-False]" -q > gpurun_out/r02_racecheck_full.log 2>&1
grep -v "Host Frame" gpurun_out/r02_racecheck_full.log | head -80 > gpurun_out/r02_racecheck_head.log
ls -la gpurun_out/ | tail -20
