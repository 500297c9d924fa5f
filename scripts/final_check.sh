#!/bin/bash
# Round-end check on one B200 (through gpurun): whole GPU suite, smoke, default bench line -> gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
e = d.get("extras", {})
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["achieved"], d["e2e"]["value"], d["clocks"])
for k in ("llama3_8b_24sparse_bfp12_weight_cast", "llama3_70b_sbfp12_weight_cast", "llama3_70b_sbfp12_packed_storage"):
    print(k, e.get(k))
o = e.get("opt125m_basic_forward", {})
print({k: (v if not isinstance(v, dict) else {a: v[a] for a in v if "overhead" in a or a.startswith("ms_")}) for k, v in o.items() if k != "config"})
P
