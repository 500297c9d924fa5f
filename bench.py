#!/usr/bin/env python
"""bench.py -- BFP cast throughput on B200 (BASELINE.json metric: "BFP cast GB/s & % of HBM peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): the standalone
block-size-64 cast sweep point  n = 2^28 elements as [65536, 4096]:
    BFP16 = BFP[8|8]{64}(SN) and BFP12 = BFP[4|8]{64}(SN), on fp32 and on bf16 tensors.
One "step" = those four casts over the batch.  Algorithmic bytes per cast = read the input once
+ write the output once = 2 * sizeof(dtype) * n: 6 GiB per step.  Each
cast touches 1-2 GiB, far above the 126 MB L2, so every step streams from HBM (no L2 flush
needed; stated in config.l2).

value     whole-job GB/s with inputs resident in HBM (device timed with CUDA events around the
          K steps, barrier + synchronize on both sides, max over ranks).
e2e       the same metric through the C ABI host entry (dmxq_cast_chain_host): pinned HOST
          buffers in, host buffers out, H2D + kernel + D2H inside the timed region.
roofline  the dominant kernel (chain_rows_kernel<float,float,flat,bfp>; the two fp32 casts are
          2/3 of the step's bytes): algorithmic bytes per launch / its mean duration measured
          with CUDA events around every launch of it inside the timed region.
cpu_baseline  the reference's own CPU path on a bounded sample, on the host cores of this box.
--impl reference  times that CPU path alone (rank 0 only), same metric / unit / config.

N > 1 (torchrun): every rank runs the same workload on its own GPU (independent tensors, no
data-path collective) -> weak scaling; value = bytes of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELEMS = 1 << 28
COLS = 4096
FORMATS = [("BFP16_64", "BFP[8|8]{64}(SN)", 8), ("BFP12_64", "BFP[4|8]{64}(SN)", 4)]
CPU_SAMPLE_ELEMS = 1 << 22


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the model-level extras (OPT-125m forward, sharded weight casts)")
    return ap.parse_args()


def config_dict(n_gpus):
    return {
        "workload": "configs[1]: standalone BFP16_64 + BFP12_64 cast of fp32 and bf16 [65536,4096] (2^28 elements), block 64 along the last dim",
        "formats": [f[1] for f in FORMATS], "dtypes": ["fp32", "bf16"], "elements_per_cast": N_ELEMS,
        "bytes_per_step": bytes_per_step(), "l2": "inputs (1-2 GiB per cast) exceed the 126 MB L2; no flush needed",
        "parallelism": f"independent replicas x{n_gpus} (no collective on the data path)",
    }


def bytes_per_step():
    return sum(2 * es * N_ELEMS for es in (4, 2)) * len(FORMATS)


# ------------------------------------------------------------------------------------------ CPU path
def _make_rows(n, seed=0):
    import torch

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n // COLS, COLS, generator=g)
    return x * torch.pow(2.0, torch.randint(-8, 9, (n // COLS, 1), generator=g).float())


def _reference_cpu_cast():
    """-> (callable(x_fp32_or_bf16_tensor, precision) -> tensor, kind, cores).

    kind "reference": the reference's compiled quant_cpu kernels (oracle/_ref/ref_quant_cpu.so, built
    from /root/reference by oracle/build_ref.py) driven by the restated python loop of
    BlockFloatingPoint.cast (S/numerical/format.py:322-341) + CastTo.forward's dtype round trip
    (S/numerical/cast.py:262,306) -- i.e. exactly what the reference executes for a CPU tensor.
    kind "port": oracle/dmxq_oracle.c when the reference binary is not available."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        import build_ref

        ref = build_ref.load("ref_quant_cpu")
    except Exception:
        ref = None
    if ref is not None:
        def cast(x, precision):
            dt = x.dtype
            _x = x.float().transpose(-1, -1)
            shp = _x.shape
            chunks = torch.split(_x.reshape(-1, shp[-1]), 64, dim=-1)
            out = [ref.block_quantize_nearest(c.contiguous(), precision, 0, True) for c in chunks]
            return torch.cat(out, dim=-1).reshape(shp).to(dt)

        return cast, "reference", cores
    import oracle as O

    def cast(x, precision):
        dt = x.dtype
        y = O.bfp_cast(x.float().numpy(), -1, 64, precision)
        return torch.from_numpy(y).to(dt)

    return cast, "port", 1


def cpu_pass(cast, xs):
    """one bounded-sample pass of the workload; returns (seconds, algorithmic bytes)"""
    t0 = time.perf_counter()
    nbytes = 0
    for x in xs:
        for _, _, prec in FORMATS:
            cast(x, prec)
            nbytes += 2 * x.element_size() * x.numel()
    return time.perf_counter() - t0, nbytes


def run_reference_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cast, kind, cores = _reference_cpu_cast()
    x32 = _make_rows(CPU_SAMPLE_ELEMS)
    xs = [x32, x32.to(torch.bfloat16)]
    for _ in range(args.warmup):
        cpu_pass(cast, xs)
    t = 0.0
    nbytes = 0
    for _ in range(args.steps):
        dt, nb = cpu_pass(cast, xs)
        t += dt
        nbytes += nb
    gbs = nbytes / t / 1e9
    sample = f"each step = the 4 casts on a bounded sample of n=2^{CPU_SAMPLE_ELEMS.bit_length() - 1} elements per tensor ([{CPU_SAMPLE_ELEMS // COLS},{COLS}])"
    line = {
        "impl": "reference", "metric": "BFP cast GB/s", "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for t, ln in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [s.strip() for s in ln.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ extras
LLAMA = {
    "8b": dict(layers=32, d=4096, kv=1024, ffn=14336, vocab=128256),
    "70b": dict(layers=80, d=8192, kv=1024, ffn=28672, vocab=128256),
}


def llama_shapes(name, layers=None):
    c = LLAMA[name]
    shapes = {}
    for i in range(layers if layers is not None else c["layers"]):
        for n, s in (("q", (c["d"], c["d"])), ("k", (c["kv"], c["d"])), ("v", (c["kv"], c["d"])), ("o", (c["d"], c["d"])),
                     ("gate", (c["ffn"], c["d"])), ("up", (c["ffn"], c["d"])), ("down", (c["d"], c["ffn"]))):
            shapes[f"layers.{i}.{n}"] = s
    shapes["lm_head"] = (c["vocab"], c["d"])
    return shapes


def extra_weight_cast(dev, rank, world, dist, model, stages_fn, layers, with_stats, dtype, packed=False):
    """Whole-model weight cast sharded by parameter / row range: every rank
    materialises its shards (random, shard-local), optionally reduces per-tensor amax with ONE
    batched all-reduce, and casts each shard with one fused kernel.  Timed on the device, max
    over ranks; value = algorithmic bytes of all ranks / time.  `packed`: write the packed SBFP storage
    (nibble mantissas + one scaler byte per block, dmxq_sbfp_pack) instead of the dequantised tensor."""
    import torch

    from dmx_compressor_b200 import parallel as P

    shapes = llama_shapes(model, layers)
    plan = P.plan_shards(shapes, world, row_align=1)
    mine = plan[rank]
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    ws = []
    for sh in mine:
        cols = shapes[sh.name][1]
        ws.append(torch.randn(sh.row1 - sh.row0, cols, device=dev, dtype=torch.float32, generator=g).mul_(0.02).to(dtype))
    from dmx_compressor_b200 import ops

    if packed:
        outs = [(torch.empty(w.shape[0], w.shape[1] // 2, device=dev, dtype=torch.uint8),
                 torch.empty(w.shape[0], w.shape[1] // 16, device=dev, dtype=torch.uint8)) for w in ws]
        inexact = torch.zeros((), device=dev, dtype=torch.int32)
    else:
        outs = [torch.empty_like(w) for w in ws]

    def run():
        if with_stats:
            # per-tensor amax: local dmxq_minmax per shard + ONE all_reduce(MAX) for the row-split tensors
            stats = P.shard_stats(plan, rank, ws)
            amax = torch.cat([torch.maximum(-mn, mx).reshape(-1) for mn, mx in stats]).cpu().tolist()  # one host sync
            for w, y, a in zip(ws, outs, amax):
                if packed:
                    ops.sbfp_pack(w, stages_fn(a)[0], out=y, inexact=inexact)
                else:
                    ops.cast_chain(w, stages_fn(a), -1, out=y)
        else:
            st = stages_fn(None)
            for w, y in zip(ws, outs):
                ops.cast_chain(w, st, -1, out=y)

    run()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if packed:
        nbytes = sum(w.numel() * w.element_size() + m.numel() + sc.numel() for w, (m, sc) in zip(ws, outs))
    else:
        nbytes = sum(2 * w.numel() * w.element_size() for w in ws)
    nelem = sum(w.numel() for w in ws)
    n_inexact = int(inexact) // 2 if packed else None  # (two runs accumulated)
    t = torch.tensor([ms, float(nbytes), float(nelem)], device=dev, dtype=torch.float64)
    if dist is not None:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_max, total, nelem = float(tm[0]), float(t[1]), float(t[2])
        ms_mean = float(t[0]) / world
    else:
        ms_max, total, ms_mean = ms, float(nbytes), ms
    del ws, outs
    torch.cuda.empty_cache()
    res = {"Gelem/s": round(nelem / (ms_max * 1e-3) / 1e9, 1)}
    if packed:
        res["blocks_not_representable"] = n_inexact
    return {**res, "GB/s": round(total / (ms_max * 1e-3) / 1e9, 1), "ms": round(ms_max, 3), "bytes": int(total), "tensors": len(shapes),
            "imbalance_max_over_mean_time": round(ms_max / ms_mean, 3), "plan_imbalance": round(P.plan_imbalance(plan, shapes), 3),
            "dtype": str(dtype).split(".")[-1], "layers": layers if layers is not None else LLAMA[model]["layers"]}


def extra_sweep(dev):
    """configs[1] in full: BFP16_64 / BFP12_64 on fp32 / bf16 tensors of 2^20 .. 2^30 elements ([n/4096, 4096]).
    Sizes whose input + output fit the 126 MB L2 are timed over 8 rotating buffer pairs so that every launch
    streams from HBM (labelled hbm_rotating); CUDA events around 20 back-to-back launches, median of 5."""
    import torch

    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format

    out = []
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        for e in (20, 22, 24, 26, 28, 30):
            n = 1 << e
            footprint = 2 * n * es
            nbuf = 8 if footprint * 2 < (1 << 30) else 1
            xs = [torch.randn(n // COLS, COLS, device=dev).to(dt) for _ in range(nbuf)]
            ys = [torch.empty_like(x) for x in xs]
            for name, sh, _ in FORMATS:
                st = [Format.from_shorthand(sh).stage()]
                reps = 20
                for i in range(3):
                    ops.cast_chain(xs[i % nbuf], st, -1, out=ys[i % nbuf])
                ts = []
                for _ in range(5):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    a.record()
                    for i in range(reps):
                        ops.cast_chain(xs[i % nbuf], st, -1, out=ys[i % nbuf])
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) / reps)
                ts.sort()
                row = {"format": name, "dtype": str(dt).split(".")[-1], "log2_elements": e, "us_per_cast": round(ts[2] * 1e3, 2),
                       "GB/s": round(footprint / ts[2] / 1e6, 1), "mode": "hbm_rotating" if nbuf > 1 else "hbm"}
                if e <= 24:
                    # below ~2^24 elements a cast is shorter than the python + launch path (~15 us): the same 20
                    # launches replayed from a CUDA graph show the device-side time
                    g = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        ops.cast_chain(xs[0], st, -1, out=ys[0])
                    torch.cuda.current_stream().wait_stream(side)
                    with torch.cuda.graph(g):
                        for i in range(reps):
                            ops.cast_chain(xs[i % nbuf], st, -1, out=ys[i % nbuf])
                    g.replay()
                    tg = []
                    for _ in range(5):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        a.record()
                        g.replay()
                        b.record()
                        torch.cuda.synchronize()
                        tg.append(a.elapsed_time(b) / reps)
                    tg.sort()
                    row["us_per_cast_cuda_graph"] = round(tg[2] * 1e3, 2)
                    row["GB/s_cuda_graph"] = round(footprint / tg[2] / 1e6, 1)
                    del g
                out.append(row)
            del xs, ys
            torch.cuda.empty_cache()
    return out


def extra_opt125m(dev):
    """BASELINE config #3: OPT-125m-shaped random-init stack, batch 8 x seq 2048 forward, BASIC rule set.
    tokens/s for the unquantised torch twin, the drop-in BASIC path, and BASIC with cast elision."""
    import torch

    from dmx_compressor_b200 import _lib, elide, opt

    res = {}
    B, S = 8, 2048

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for dt in (torch.float32, torch.bfloat16):
        q, p = opt.build_pair(device=dev, dtype=dt)
        ids = torch.randint(0, 50272, (B, S), device=dev)

        with torch.no_grad():
            t_plain = timeit(lambda: p(ids))
            n0 = _lib.launch_count()
            y1 = q(ids)
            n1 = _lib.launch_count()
            t_basic = timeit(lambda: q(ids))
            with elide.enabled():
                y2 = q(ids)
                n2 = _lib.launch_count()
                q(ids)
                n3 = _lib.launch_count()
                t_el = timeit(lambda: q(ids))
        res[str(dt).split(".")[-1]] = {
            "tokens_per_s_unquantised": round(B * S / t_plain * 1e3), "tokens_per_s_basic": round(B * S / t_basic * 1e3),
            "tokens_per_s_basic_elided": round(B * S / t_el * 1e3), "ms_unquantised": round(t_plain, 2), "ms_basic": round(t_basic, 2),
            "ms_basic_elided": round(t_el, 2), "cast_overhead_basic": round((t_basic - t_plain) / t_plain, 3),
            "cast_overhead_basic_elided": round((t_el - t_plain) / t_plain, 3), "dmxq_launches_basic": n1 - n0,
            "dmxq_launches_elided": n3 - n2, "elided_equals_dropin_bitwise": bool(torch.equal(y1, y2))}
        del q, p, y1, y2
        torch.cuda.empty_cache()
    # small-batch latency: launch-bound eagerly, so also as a captured CUDA graph
    try:
        from dmx_compressor_b200 import graph

        q, p = opt.build_pair(device=dev, dtype=torch.float32)
        ids = torch.randint(0, 50272, (4, 128), device=dev)
        with torch.no_grad():
            want = q(ids)
            t_eager = timeit(lambda: q(ids), 5)
        fq, fp = graph.capture(q, ids), graph.capture(p, ids)
        res["small_batch_4x128_fp32"] = {"ms_basic_eager": round(t_eager, 2), "ms_basic_cuda_graph": round(timeit(lambda: fq(ids), 20), 2),
                                         "ms_unquantised_cuda_graph": round(timeit(lambda: fp(ids), 20), 2),
                                         "graph_equals_eager_bitwise": bool(torch.equal(fq(ids), want))}
        del q, p, fq, fp
        torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        res["small_batch_4x128_fp32"] = {"error": repr(e)}
    res["config"] = "OPT-125m shape (12 layers, d=768, ffn=3072, 12 heads, vocab 50272), random init, batch 8 x seq 2048, config_rules.BASIC"
    return res


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import dmx_compressor_b200 as dmx  # raises if libdmxq.so is missing: no fallback
    from dmx_compressor_b200 import ops
    from dmx_compressor_b200.numerical import Format

    stages = {name: [Format.from_shorthand(sh).stage()] for name, sh, _ in FORMATS}
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x32 = torch.randn(N_ELEMS // COLS, COLS, device=dev, generator=g)
    x32 *= torch.pow(2.0, torch.randint(-8, 9, (N_ELEMS // COLS, 1), device=dev, generator=g).float())
    x16 = x32.to(torch.bfloat16)
    y32, y16 = torch.empty_like(x32), torch.empty_like(x16)
    casts = [(x32, y32, n) for n in stages] + [(x16, y16, n) for n in stages]

    def step(events=None):
        for x, y, name in casts:
            if events is not None and x.dtype == torch.float32:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.cast_chain(x, stages[name], -1, out=y)
                b.record()
                events.append((a, b))
            else:
                ops.cast_chain(x, stages[name], -1, out=y)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = dmx._lib.launch_count()
    kern_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step(kern_events)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = dmx._lib.launch_count() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = bytes_per_step() * args.steps * world / (ms * 1e-3) / 1e9
    kms = sorted(a.elapsed_time(b) for a, b in kern_events)
    k_mean = sum(kms) / len(kms)

    # ---------------- e2e: host buffers through the C ABI host entry (H2D + kernel + D2H timed)
    e2e = None
    if args.e2e_steps > 0:
        xh32 = torch.empty(x32.shape, dtype=torch.float32, pin_memory=True)
        xh32.copy_(x32)
        xh16 = torch.empty(x16.shape, dtype=torch.bfloat16, pin_memory=True)
        xh16.copy_(x16)
        yh32 = torch.empty(x32.shape, dtype=torch.float32, pin_memory=True)
        yh16 = torch.empty(x16.shape, dtype=torch.bfloat16, pin_memory=True)
        hcasts = [(xh32, yh32, n) for n in stages] + [(xh16, yh16, n) for n in stages]

        def hstep():
            for xh, yh, name in hcasts:
                ops.cast_chain_host(xh, yh, stages[name], local)

        hstep()
        barrier()
        th0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hstep()
        torch.cuda.synchronize()
        th = time.perf_counter() - th0
        if dist is not None:
            t = torch.tensor([th], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            th = float(t.item())
        # check the host path against the device path on the last cast
        ok = torch.equal(yh16.view(torch.int16)[:64], y16[:64].cpu().view(torch.int16))
        half = bytes_per_step() // 2
        e2e = {"value": round(bytes_per_step() * args.e2e_steps * world / th / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": half, "d2h_bytes_per_step": half, "steps": args.e2e_steps,
               "api": "dmxq_cast_chain_host (pinned host in/out, 3-stream chunked pipeline)", "matches_device_path": bool(ok)}
        del xh32, xh16, yh32, yh16

    # ---------------- extras: the model-level configs of BASELINE.json (not part of `value`)
    extras = {}
    if not args.no_extras:
        from dmx_compressor_b200 import parallel as P

        f12 = Format.from_shorthand("BFP[4|8]{64}(SN)").stage()
        # config #4: Llama-3-8B-shaped weights, 2:4 sparsity (score |w|) -> BFP12, sharded over the ranks
        extras["llama3_8b_24sparse_bfp12_weight_cast"] = extra_weight_cast(
            dev, rank, world, dist, "8b", lambda amax: [ops.nm_stage(2, 4), f12], None, False, torch.bfloat16)

        # config #5: Llama-3-70B-shaped weights, SBFP12_16 with the scaler bias chosen from the amax all-reduce
        def sbfp(amax):
            b = 7 if amax is None else P.sbfp_scaler_bias_from_amax(amax)
            return [Format.from_shorthand(f"SBFP<XP[4,0](CSN)><FP[0|4|4,{b}](FN)>{{16}}").stage()]

        extras["llama3_70b_sbfp12_weight_cast"] = extra_weight_cast(
            dev, rank, world, dist, "70b", sbfp, 10 * world, True, torch.bfloat16)
        extras["llama3_70b_sbfp12_weight_cast"]["note"] = "10 layers per GPU (weak scaling; 80 layers at 8 GPUs), one batched amax all-reduce"
        # the same, written as packed storage (0.5625 B per element instead of a dequantised bf16 tensor)
        extras["llama3_70b_sbfp12_packed_storage"] = extra_weight_cast(
            dev, rank, world, dist, "70b", sbfp, 10 * world, True, torch.bfloat16, packed=True)
        extras["llama3_70b_sbfp12_packed_storage"]["note"] = "dmxq_sbfp_pack: bf16 in, nibble mantissas + E4M4 scaler byte out (2.5625 B/elem algorithmic)"
        if rank == 0:
            try:
                extras["cast_sweep"] = extra_sweep(dev)
            except Exception as e:  # pragma: no cover
                extras["cast_sweep"] = {"error": repr(e)}
            try:
                extras["opt125m_basic_forward"] = extra_opt125m(dev)
            except Exception as e:  # pragma: no cover
                extras["opt125m_basic_forward"] = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    k_bytes = 8 * N_ELEMS
    achieved = k_bytes / (k_mean * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("chain_rows_kernel_f32_flat_bfp", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "chain_rows_kernel<float,float,FLAT,SPECIAL=2> (BFP16/BFP12 fp32 casts)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "frac_of_8TBps_nominal": round(achieved / 8000.0, 4), "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_launch": k_bytes, "launch_ms_mean": round(k_mean, 4), "launch_ms_median": round(kms[len(kms) // 2], 4),
                "launches_timed": len(kms)}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cast, kind, cores = _reference_cpu_cast()
        c32 = _make_rows(CPU_SAMPLE_ELEMS)
        xs = [c32, c32.to(torch.bfloat16)]
        cpu_pass(cast, xs)
        best = None
        tot = 0.0
        while tot < 10.0:
            dt, nb = cpu_pass(cast, xs)
            tot += dt
            best = dt if best is None else min(best, dt)
        cpu_baseline = {"value": round(nb / best / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": kind,
                        "sample": f"the step's 4 casts on n=2^{CPU_SAMPLE_ELEMS.bit_length() - 1} elements per tensor, best pass of ~10 s of CPU work"}

    line = {
        "metric": "BFP cast GB/s", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(world), "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "frac_of_hbm_peak": round(value / world / peak, 4), "extras": extras,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
