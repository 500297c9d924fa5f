#!/bin/bash
# Round-end check on one B200 (through gpurun): whole GPU suite, smoke, default bench line -> gpurun_out/.
# usage: bash scripts/final_check.sh [r02]
R=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${R}_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/${R}_bench_1gpu.jsonl 2> gpurun_out/${R}_bench_1gpu.err ) 2>&1 | grep real
python - "$R" <<'P'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench_1gpu.jsonl").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["achieved"], d["e2e"], d["clocks"])
print("sharded", json.dumps(d.get("sharded")))
print("reference_cuda", d.get("reference_cuda"))
print("cpu_baseline", d.get("cpu_baseline"))
for r in d.get("roofline_by_format") or []:
    print("  ", r)
for k in ("opt125m_basic_forward", "opt125m_reference_modules_plus_plugin"):
    print(k, json.dumps(d.get(k)))
P
