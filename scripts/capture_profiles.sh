#!/bin/bash
# Re-create the ncu evidence under gpurun_out/ (run on the GPU box through gpurun), summarised there with
# scripts/ncu_summarise.py; copy gpurun_out/<r>_*.csv into profiles/ afterwards.
# usage: bash scripts/capture_profiles.sh r01
set -x
R=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-extras --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv $B --steps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_rows -s 12 -c 4 -o gpurun_out/prof_${R}_bench -f $B --steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:chain_|bfp_cols16|histc_kernel|minmax_flat|minmax_cols|bfp_pack|bfp_unpack|fixed_chan' -c 27 -o gpurun_out/prof_${R}_other -f python scripts/ncu_two.py > /dev/null 2>&1
python scripts/ncu_summarise.py gpurun_out/prof_${R}_bench.ncu-rep gpurun_out/${R}_ncu_full_bench_step.csv
python scripts/ncu_summarise.py gpurun_out/prof_${R}_other.ncu-rep gpurun_out/${R}_ncu_full_other_kernels.csv
rm -f gpurun_out/prof_${R}_other.ncu-rep   # (64 MiB return limit; the bench report is kept for --page source reading)
ls -la gpurun_out/
