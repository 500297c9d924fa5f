#!/bin/bash
# Source-level instruction counters of one rows-kernel launch per case (development aid; run through gpurun).
# usage: bash scripts/ncu_src.sh sbfp_bf16 nm_bfp_bf16 ...   -> gpurun_out/<case>.src.csv + the dynamic opcode mix on stdout
mkdir -p gpurun_out
for k in "$@"; do
  ncu --section SourceCounters --clock-control none --import-source on -k "regex:${NCU_KERNEL:-chain_rows}" -s 2 -c 1 -o gpurun_out/src_$k -f python scripts/ncu_one.py $k > /dev/null 2>&1
  ncu -i gpurun_out/src_$k.ncu-rep --page source --csv > gpurun_out/$k.src.csv 2>/dev/null
  rm -f gpurun_out/src_$k.ncu-rep
  echo "== $k"; python scripts/ncu_mix.py gpurun_out/$k.src.csv | head -28
done
