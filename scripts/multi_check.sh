#!/bin/bash
# N-GPU check (through `gpurun --gpus N`): the torchrun parity test of the sharded weight cast, the bench line at N GPUs
# (as the driver launches it), the host topology the e2e path runs on.   usage: bash scripts/multi_check.sh N [r02]
N=${1:-2}
R=${2:-r02}
mkdir -p gpurun_out
( nvidia-smi topo -m; echo; lscpu | grep -i -E "^CPU\(s\)|numa|socket|model name"; echo; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; which numactl ) > gpurun_out/${R}_topology_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_parallel_gputests_${N}gpu.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 100 --warmup 5 --no-details \
    > gpurun_out/${R}_bench_${N}gpu.jsonl 2> gpurun_out/${R}_bench_${N}gpu.err ) 2>&1 | grep real
python - "$R" "$N" <<'P'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench_{sys.argv[2]}gpu.jsonl").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")}, d["e2e"], d["clocks"])
print("sharded", json.dumps(d.get("sharded")))
P
tail -3 gpurun_out/${R}_bench_${N}gpu.err
