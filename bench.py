#!/usr/bin/env python
"""Benchmark of the HybridFusion hot path (BASELINE.json metric: windows/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N: BASELINE.json configs[1] per GPU — HybridFusion train step
(forward + CE(label smoothing 0.05) + backward + grad-norm clip + AdamW),
bf16 tensor-core path, 4096 PAMAP2-shaped windows per GPU, dropout 0.1 (weak
scaling: global batch 4096*N, gradients all-reduced over NCCL).  One JSON line
on stdout (rank 0).  `value` = windows/s with inputs resident in HBM, `e2e` =
the same step fed from pinned host buffers through the public engine API.

`--impl reference`: the reference's own CPU path.  The reference is Python and
cannot travel to the GPU box, so this arm times the CPU oracle port
(oracle/fusion_oracle.py, pinned against the reference by tests/golden) with
all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "multimodal-sensor-fusion-with-attention-rajeevatla_b200"

DIMS = {"imu_hand": 128, "imu_chest": 128, "imu_ankle": 128, "heart_rate": 128}
HIDDEN, HEADS, CLASSES, BATCH, DROPOUT, SMOOTHING = 256, 4, 25, 4096, 0.1, 0.05
# live-path FLOPs per window: forward 3 553 792, train step 3x
def _flop_fwd():
    m = len(DIMS)
    return sum(2 * d * HIDDEN for d in DIMS.values()) + 4 * HIDDEN * HIDDEN * m * (m - 1) + 2 * HIDDEN * m \
        + 2 * HIDDEN * HIDDEN + 2 * HIDDEN * CLASSES


FLOP_FWD = _flop_fwd()
FLOP_TRAIN = 3 * FLOP_FWD
SHAPE = "pamap2"


def use_scaled_shape():
    """BASELINE configs[4] (SURVEY config 5): 8 modalities of width 256 (2 video + 6 IMU), hidden 512, 8 heads,
    11 classes (config/base.yaml:57-65).  Outside the fused kernels (M > 4, H > 256): grouped tcgen05 GEMM launches."""
    global DIMS, HIDDEN, HEADS, CLASSES, FLOP_FWD, FLOP_TRAIN, SHAPE
    DIMS = {f"video_{i}": 256 for i in range(2)}
    DIMS.update({f"imu_{i}": 256 for i in range(6)})
    HIDDEN, HEADS, CLASSES = 512, 8, 11
    FLOP_FWD = _flop_fwd()
    FLOP_TRAIN = 3 * FLOP_FWD
    SHAPE = "scaled"
METRIC = "windows/sec HybridFusion fwd+bwd & inference at 1/2/4/8 B200; % of HBM/TC roofline"


class HostPipeline:
    """Two device slots for a list of input tensors fed from pinned host memory on a copy stream: the copy of batch
    i+1 overlaps the pass over batch i (what FusionEngine.train_stream does for the fusion step).  next() returns the
    device tensors of the batch to process now (the compute stream already waits for its copy); release() marks them
    consumed, so that the slot may be overwritten."""

    def __init__(self, torch, host_tensors, device):
        self.torch, self.host = torch, host_tensors
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [[torch.empty(t.shape, dtype=t.dtype, device=device) for t in host_tensors] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self._issue(0)

    def _issue(self, k):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[k])
            for dst, src in zip(self.slots[k], self.host):
                dst.copy_(src, non_blocking=True)
            self.ready[k].record(self.copy_stream)

    def next(self):
        k = self.i & 1
        self._issue(k ^ 1)
        self.torch.cuda.current_stream().wait_event(self.ready[k])
        return self.slots[k]

    def release(self):
        self.free[self.i & 1].record(self.torch.cuda.current_stream())
        self.i += 1


def _lstm_traffic(key, steps):
    """DRAM bytes of one recurrence launch from the ncu capture (profiles/traffic.json: per time step), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["lstm_kernels_dram_bytes_per_time_step"][key] * steps
    except Exception:  # noqa: BLE001 - optional ncu-derived figure
        return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "tflops_burst": p.get("bf16_tflops"),
                "src": "measured (sustained)"}
    except Exception:  # noqa: BLE001 - profiling guide's stated fallback
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(torch, seed, batch, device=None, pin=False):
    """x ~ N(0,1), mask ~ Bernoulli(0.9) with all-missing rows
    re-drawn to one modality, labels uniform."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(batch, d, generator=g) for d in DIMS.values()]
    mask = (torch.rand(batch, len(DIMS), generator=g) < 0.9).float()
    dead = mask.sum(1) == 0
    mask[dead, torch.randint(0, len(DIMS), (int(dead.sum()),), generator=g)] = 1.0
    labels = torch.randint(0, CLASSES, (batch,), generator=g)
    if device is not None:
        return [f.to(device) for f in feats], mask.to(device), labels.to(device)
    if pin:
        return [f.pin_memory() for f in feats], mask.pin_memory(), labels.pin_memory()
    return feats, mask, labels


def oracle_cpu_throughput(torch, budget_s: float, warmup: int = 1, steps=None):
    """Reference arithmetic on the host: oracle forward + autograd backward + AdamW (CPU port)."""
    from oracle import fusion_oracle
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    fusion = importlib.import_module("fusion")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    names = list(DIMS)
    feats, mask, labels = synthetic_batch(torch, 1234, BATCH)
    xs = dict(zip(names, feats))
    moments = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()}
    g = torch.Generator().manual_seed(7)
    keep = 1.0 - DROPOUT

    def draw(shape):
        return (torch.rand(shape, generator=g) < keep).float() / keep

    # dropout masks are drawn OUTSIDE the timed region (a pool of 4 sets, used in turn): the reference's
    # bernoulli_ is one pass per site, three passes of torch.rand/compare/divide here would handicap the baseline
    pool = [{"input": {m: draw((BATCH, d)) for m, d in DIMS.items()},
             "proj": {m: draw((BATCH, HIDDEN)) for m in names},
             "attn": {f"{q}_to_{k}": draw((BATCH, HEADS, 1, 1)) for q in names for k in names if q != k},
             "cls": draw((BATCH, HIDDEN))} for _ in range(4)]

    def one_step(step):
        drops = pool[step % len(pool)]
        logits, _ = fusion_oracle.hybrid_fusion_forward(sd, names, HEADS, xs, mask, drops=drops)
        loss = fusion_oracle.cross_entropy_label_smoothing(logits, labels, SMOOTHING)
        grads = torch.autograd.grad(loss, list(sd.values()))
        grads = [gr.clone() for gr in grads]
        fusion_oracle.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            for (k, p), gr in zip(sd.items(), grads):
                fusion_oracle.adamw_step(p, gr, moments[k][0], moments[k][1], step)
        return float(loss.detach())

    for i in range(warmup):
        one_step(i + 1)
    times, i = [], warmup
    t_end = time.perf_counter() + budget_s
    while (steps is None and time.perf_counter() < t_end and len(times) < 200) or \
            (steps is not None and len(times) < steps):
        t0 = time.perf_counter()
        one_step(i + 1)
        times.append(time.perf_counter() - t0)
        i += 1
    med = statistics.median(times)
    return {"value": BATCH / med, "unit": "windows/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} train steps of {BATCH} windows (oracle fp32 + autograd + AdamW, dropout masks "
                      f"pre-drawn outside the timed region), median; a reported baseline, not the target",
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    res = oracle_cpu_throughput(torch, budget_s=1e9, warmup=max(1, min(args.warmup, 2)),
                                steps=max(1, min(args.steps, 20)))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Python and cannot travel to the GPU box: CPU oracle port timed on host cores",
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    what = ("HybridFusion train step (BASELINE configs[1]): fwd + CE(ls 0.05) + bwd + clip + AdamW" if SHAPE == "pamap2" else
            "HybridFusion train step, scaled variant (BASELINE configs[4]: 8 modalities, d_model 512, 8 heads): fwd + "
            "CE(ls 0.05) + bwd + clip + AdamW")
    return {"workload": what,
            "per_gpu_batch": BATCH, "global_batch": BATCH * n, "modalities": len(DIMS),
            "feature_dim": next(iter(DIMS.values())),
            "hidden": HIDDEN, "heads": HEADS, "classes": CLASSES, "dropout": DROPOUT,
            "parallelism": f"dp{n} (batch-sharded, replicated parameters, gradient exchange per step)" if n > 1
                           else "single GPU"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    t_start = time.perf_counter()

    def mark(what):
        if os.environ.get("MSF_BENCH_TRACE"):
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:6.1f}s] {what}", file=sys.stderr, flush=True)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one JSON line: NCCL / library chatter written to fd 1 goes to stderr instead
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mark("process group ready")
    pkg = importlib.import_module(PKG)
    engine_mod = importlib.import_module(PKG + ".engine")
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    fusion = importlib.import_module("fusion")
    lib = pkg.lib()

    precision = args.precision
    torch.manual_seed(0)  # identical replicas on every rank
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT)
    # host-facing leg: features cross PCIe as bf16 by default (the tensor-core path rounds them to bf16 anyway:
    # msf_fusion_call.x_bf16); the resident ring below stays fp32
    host_bf16 = args.host_dtype == "bf16" and precision == "bf16"
    if host_bf16:   # bf16 features are read by the fused projection kernel only (msf_fusion_call.x_bf16)
        ops_mod = importlib.import_module(PKG + ".ops")
        host_bf16 = ops_mod.layer_norm_fused(model._plan(), pkg.native.MSF_PREC_BF16)
    eng = engine_mod.FusionEngine(model, BATCH, precision=precision, label_smoothing=SMOOTHING,
                                  max_grad_norm=1.0, seed=1234, use_graph=not args.no_graph,
                                  feature_dtype=torch.bfloat16 if host_bf16 else torch.float32)

    # ring of resident batches: 24 x 8.5 MB = 204 MB of inputs > 126 MB L2, so no step finds its inputs in L2
    ring_n = 24
    ring = [synthetic_batch(torch, 1000 + 97 * rank + i, BATCH, device=dev) for i in range(ring_n)]
    in_bytes = sum(t.numel() * t.element_size() for t in ring[0][0]) + ring[0][1].numel() * 4 + ring[0][2].numel() * 8
    host = [synthetic_batch(torch, 5000 + 97 * rank + i, BATCH, pin=True) for i in range(4)]
    if host_bf16:
        host = [([t.to(torch.bfloat16).pin_memory() for t in f], m, y) for f, m, y in host]
    host_bytes = sum(t.numel() * t.element_size() for t in host[0][0]) + host[0][1].numel() * 4 + host[0][2].numel() * 8
    if args.packed_host:   # the same batches in FusionEngine.pinned_batch() buffers: one transfer per batch
        packed = []
        for f, m, y in host:
            pf, pm, py = eng.pinned_batch()
            for dst, src in zip(pf + [pm, py], f + [m, y]):
                dst.copy_(src)
            packed.append((pf, pm, py))
        host = packed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, min_seconds=0.5, max_repeats=400):
        """`steps` calls of fn between two CUDA events on the launching stream (barrier + synchronize on both
        sides, max over ranks), repeated until the timed regions add up to >= min_seconds so the clock sampler
        sees the load; returns (median ms per call, repeats)."""
        for i in range(warmup):
            fn(i)
        samples, total, k = [], 0.0, warmup
        while True:
            barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            for i in range(steps):
                fn(k + i)
            stop.record()
            barrier()
            k += steps
            ms = torch.tensor([start.elapsed_time(stop)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
            samples.append(float(ms) / steps)
            total += float(ms)
            if total >= min_seconds * 1e3 or len(samples) >= max_repeats:
                break
        return statistics.median(samples), len(samples)

    # every ring batch is an input slot of its own (its own captured graph): the step reads it in place
    ring_slots = [eng.add_resident_batch(*b) for b in ring] if not args.no_graph else None

    def resident_step(i):
        if ring_slots is not None:
            eng.train_step_slot(ring_slots[i % ring_n])
        else:
            f, m, y = ring[i % ring_n]
            eng.load_batch(f, m, y)
            eng.train_step_resident()

    losses = []

    def e2e_loop(first, n):
        # public host-facing API: every batch is copied host(pinned)->device and every step's loss is read
        # back device->host inside the timed region; train_stream overlaps batch i+1's copies with step i
        for loss in eng.train_stream(host[i % len(host)] for i in range(first, first + n)):
            losses.append(loss)

    def timed_e2e(steps, warmup, min_seconds=0.25, max_repeats=100):
        e2e_loop(0, warmup)
        samples, total, first = [], 0.0, warmup
        while True:
            barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            e2e_loop(first, steps)
            stop.record()
            barrier()
            first += steps
            ms = torch.tensor([start.elapsed_time(stop)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            samples.append(float(ms) / steps)
            total += float(ms)
            if total >= min_seconds * 1e3 or len(samples) >= max_repeats:
                break
        return statistics.median(samples), len(samples)

    mark("engine and batches built")
    warm = max(3, args.warmup)
    before = lib.msf_launch_count()
    for i in range(ring_n if ring_slots is not None else 1):   # capture every slot's graph outside the timed region
        resident_step(i)
    torch.cuda.synchronize()
    mark("first step (graph captured)")
    per_step_launches = eng_launches(lib, before, eng)
    mark("launch count taken")

    # Steps per graph launch (FusionEngine.train_slots): 8, so the gap between two graph launches is paid once per
    # eight steps, at every N (the same way on one GPU and under data parallelism).  Still exactly args.steps
    # optimizer steps inside every timed region, each with the full work of a single step; what does not fill a
    # group runs as single-step launches.
    G = args.steps_per_graph if args.steps_per_graph > 0 else (8 if args.steps >= 16 else 1)
    if ring_slots is None or ring_n % G:
        G = 1   # short runs (profiler passes with a handful of steps) and eager mode keep one step per launch
    if G > 1:
        def resident_group(j):
            first = (j * G) % ring_n
            eng.train_slots([ring_slots[first + t] for t in range(G)])

        for j in range(ring_n // G):   # capture outside the timed region
            resident_group(j)
        torch.cuda.synchronize()
    warm = max(3, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if G > 1:
        groups, rest = divmod(args.steps, G)
        calls = groups + rest   # launches per timed region: `groups` graph launches of G steps + `rest` single steps

        def launch(j):
            if j % calls < groups:
                resident_group(j)
            else:
                resident_step(j)

        ms_call, repeats = timed(launch, calls, -(-warm // G))
        ms = ms_call * calls / args.steps
    else:
        ms, repeats = timed(resident_step, args.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    mark("resident loop timed")
    ms_e2e, repeats_e2e = timed_e2e(max(5, min(args.steps, 100)), 3)

    mark("e2e loop timed")
    # every rank runs the profiled steps: the eager step contains the gradient all-reduce
    kern = profile_dominant_kernel(torch, pkg, eng, ring, ring_n)
    mark("kernel profile done")

    # Data-parallel correctness, outside the timed region: every rank's parameter arena (fp32 master and the bf16
    # compute copy) must be bit-identical after the run.  Compared as 64-bit checksums (max == min over ranks).
    replicas_identical = None
    if world > 1:
        eng.gather_parameters()   # sharded step: collect every weight tile from its owner first

        def checksum(t):
            v = t.view(torch.int32).to(torch.int64)
            return torch.stack([v.sum(), (v * torch.arange(1, v.numel() + 1, device=dev) % 1000003).sum()])
        cs = checksum(eng.arena)
        if eng.arena_bf16 is not None:
            raw = eng.arena_bf16.view(torch.uint8)
            cs = torch.cat([cs, checksum(raw[:raw.numel() // 4 * 4])])
        hi, lo = cs.clone(), cs.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        replicas_identical = bool(torch.equal(hi, lo))
    comm_used = eng.comm

    # BASELINE configs[3] (SURVEY "Config 4"): STRONG scaling, global batch 32768 split over the N ranks
    strong = None
    if not args.no_strong and 32768 % world == 0:
        per = 32768 // world
        torch.manual_seed(0)
        model_s = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT)
        eng_s = engine_mod.FusionEngine(model_s, per, precision=precision, label_smoothing=SMOOTHING,
                                        max_grad_norm=1.0, seed=1234, use_graph=not args.no_graph)
        n_ring = max(2, min(8, (160 << 20) // (per * 2072) + 1))     # > 126 MiB L2 of inputs in rotation
        ring_s = [synthetic_batch(torch, 3000 + 97 * rank + i, per, device=dev) for i in range(n_ring)]
        slots_s = [eng_s.add_resident_batch(*b) for b in ring_s] if not args.no_graph else None

        def strong_step(i):
            if slots_s is not None:
                eng_s.train_step_slot(slots_s[i % n_ring])
            else:
                eng_s.load_batch(*ring_s[i % n_ring])
                eng_s.train_step_resident()

        for i in range(n_ring):
            strong_step(i)
        torch.cuda.synchronize()
        ms_s, rep_s = timed(strong_step, max(4, min(args.steps, 20)), 3, min_seconds=0.2)
        strong = {"global_batch": 32768, "per_gpu_batch": per, "ms_per_step": ms_s, "value": 32768 / (ms_s * 1e-3),
                  "unit": "windows/s", "scaling": "strong", "steps": max(4, min(args.steps, 20)), "repeats": rep_s,
                  "collective": eng_s.comm}
        mark("strong-scaling record timed")
        eng = eng_s   # finish() drops the graphs of the live engine

    def finish():
        """Leave without tearing the NCCL communicator down: destroy_process_group() blocks while CUDA
        graphs that captured collectives on it are alive, so drop the graphs, meet at a barrier and exit."""
        if world > 1:
            eng._train_graph = eng._infer_graph = None
            eng._train_graphs = [None] * len(eng._train_graphs)
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return
    peaks = _peaks()
    value = BATCH * world / (ms * 1e-3)
    step_tflops = FLOP_TRAIN * BATCH / (ms * 1e-3) / 1e12  # per GPU, whole fused step
    line = {
        "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(world),
        "run": {"l2": f"inputs rotate through a ring of {ring_n} resident batches "
                      f"({ring_n * in_bytes / 2**20:.0f} MiB > 126 MiB L2), read in place",
                "cuda_graph": not args.no_graph, "steps_per_graph_launch": G, "warmup_done": max(warm, -(-warm // G) * G),
                "repeats": repeats, "timed_region_ms": ms * args.steps * repeats,
                "ms_per_step_is": "median over `repeats` timed regions of exactly `steps` steps each",
                "collective": {"none": "none (single GPU)",
                               "zshard": "sharded optimizer over NVLink peer memory (opt_pack.cu: dpz_reduce_kernel + "
                                         "opt_pack_kernel<.., true>): push-based reduce-scatter of the gradient tiles to "
                                         "their owners, owner-side clip + AdamW, bf16 weights pushed to all ranks; no NCCL "
                                         "on the data path",
                               "p2p": "fused NVLink peer-memory reduce-scatter + all-gather kernels (dp_optim.cu), no NCCL on the data path",
                               "nccl": "ncclAllReduce of the flat fp32 gradient arena"}[comm_used]},
        "clocks": clocks,
        "e2e": {"value": BATCH * world / (ms_e2e * 1e-3), "unit": "windows/s", "h2d_bytes_per_step": host_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "repeats": repeats_e2e,
                "host_feature_dtype": "bf16" if host_bf16 else "f32",
                "host_batch": "one pinned allocation per batch, one transfer" if args.packed_host
                              else "one pinned tensor per modality + mask + labels, one copy each",
                "api": "FusionEngine.train_stream(host batches): 2-slot pipeline, H2D of batch i+1 overlaps step i"},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "roofline": roofline_entry(kern, peaks, step_tflops),
        "final_loss": losses[-1] if losses else None,
    }
    if replicas_identical is not None:
        line["replicas_identical"] = replicas_identical
    if strong is not None:
        line["strong_32768"] = strong
    if world == 1 and not args.no_cpu_baseline:
        cpu = oracle_cpu_throughput(torch, budget_s=12.0)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    finish()


FLOP_SWEEP_MASK_AWARE = 16854528 / 15.0   # live FLOPs per window-evaluation when absent modalities
                                          # (their projections, pair GEMMs, query rows) are skipped: 1 123 635


def _timed_repeats(torch, fn, steps, warmup, min_seconds=0.5, max_repeats=200):
    """`steps` calls of fn between CUDA events, repeated until >= min_seconds are covered; median ms per call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    samples, total = [], 0.0
    while True:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop)
        samples.append(ms / steps)
        total += ms
        if total >= min_seconds * 1e3 or len(samples) >= max_repeats:
            break
    return statistics.median(samples), len(samples)


def run_infer_sweep(args):
    """BASELINE configs[2]: inference over all 15 missing-modality subsets (eval.py:342-348 order), batch
    65536 per subset, single GPU; logits -> softmax -> (conf, pred) -> 15-bin ECE statistics per subset.
    Throughput = 15 * B window-evaluations / time (inputs resident; one CUDA graph per subset mask).
    `e2e`: the same sweep fed from pinned host features (copied once per sweep) with every subset's predictions and
    bin statistics read back; `cpu_baseline`: the oracle's forward + softmax/argmax + binning on the host cores."""
    import itertools
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    pkg = importlib.import_module(PKG)
    engine_mod = importlib.import_module(PKG + ".engine")
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    fusion = importlib.import_module("fusion")
    B = 65536
    torch.manual_seed(0)
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT).eval()
    eng = engine_mod.FusionEngine(model, B, precision=args.precision, use_graph=not args.no_graph)
    host_feats, _, host_labels = synthetic_batch(torch, 1234, B, pin=True)
    feats, labels = [f.to(dev) for f in host_feats], host_labels.to(dev)
    names = list(DIMS)
    subsets = [c for r in range(1, len(names) + 1) for c in itertools.combinations(range(len(names)), r)]
    masks = []
    for sub in subsets:
        m = torch.zeros(B, len(names), device=dev)
        m[:, list(sub)] = 1.0
        masks.append(m)
    edges = torch.linspace(0, 1, 16).double().tolist()
    stats = torch.zeros(len(subsets), 3, 15, dtype=torch.int64, device=dev)
    host_pred = torch.zeros(len(subsets), B, dtype=torch.int64).pin_memory()
    host_stats = torch.zeros(len(subsets), 3, 15, dtype=torch.int64).pin_memory()

    folded = not args.no_mask_hint and not args.no_fold

    def sweep(src_feats=feats, readback=False):
        if folded:   # projections shared over the sweep, one folded pair GEMM + head kernel per subset
            def after(i, sub):
                eng.ece_bins(labels, edges, out=stats[i])
                if readback:
                    host_pred[i].copy_(eng.pred, non_blocking=True)
            eng.infer_sweep(src_feats, subsets, on_subset=after)
            if readback:
                host_stats.copy_(stats, non_blocking=True)
            return
        eng.load_batch(src_feats, masks[0], labels)  # the batch is copied once per sweep, the mask per subset
        for i, sub in enumerate(subsets):
            if args.no_mask_hint:
                eng.mask.copy_(masks[i], non_blocking=True)
                eng.infer_resident()
            else:
                eng.infer_subset(None, sub)          # uniform mask: absent modalities' work is skipped
            eng.ece_bins(labels, edges, out=stats[i])
            if readback:
                host_pred[i].copy_(eng.pred, non_blocking=True)
        if readback:
            host_stats.copy_(stats, non_blocking=True)

    steps = max(1, min(args.steps, 50))
    sampler = ClockSampler(0)
    sampler.start()
    stats.zero_()
    ms, repeats = _timed_repeats(torch, sweep, steps, max(3, args.warmup))
    clocks = sampler.stop()
    total = int(stats[:, 0].sum().item())
    stats.zero_()
    pipe = HostPipeline(torch, list(host_feats), dev)   # the copy of the next sweep's batch overlaps this sweep

    def e2e_sweep():
        sweep(pipe.next(), True)
        pipe.release()

    ms_e2e, rep_e2e = _timed_repeats(torch, e2e_sweep, max(3, min(steps, 10)), 2, min_seconds=0.3)
    del pipe
    evals = len(subsets) * B
    peaks = _peaks()
    per_eval = FLOP_FWD if args.no_mask_hint else FLOP_SWEEP_MASK_AWARE
    tf = per_eval * evals / (ms * 1e-3) / 1e12
    h2d = sum(f.numel() * 4 for f in host_feats)
    line = {"metric": METRIC, "value": evals / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "HybridFusion inference mask sweep (BASELINE configs[2]): 15 subsets x 65536 "
                                   "windows, fwd + softmax/argmax + ECE binning", "batch": B, "subsets": len(subsets),
                       "mask_hint": not args.no_mask_hint,
                       "path": ("FusionEngine.infer_sweep: projections computed once per sweep, value_proj -> out_proj "
                                "of every attention module folded into one matrix (msf_fusion_infer_folded)") if folded
                               else ("infer_subset per subset (chained pair kernels, absent modalities skipped)"
                                     if not args.no_mask_hint else "dense per-row-mask path per subset")},
            "run": {"l2": "inputs 34 MB per subset, workspace 1.7 GB >> 126 MiB L2", "repeats": repeats,
                    "timed_region_ms": ms * steps * repeats, "a_step_is": "one sweep = 15 x 65536 window-evaluations"},
            "clocks": clocks,
            "e2e": {"value": evals / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_pred.numel() * 8 + host_stats.numel() * 8,
                    "repeats": rep_e2e,
                    "api": "pinned host features copied once per sweep (2-slot pipeline on a copy stream: the copy of the "
                           "next sweep's batch overlaps this sweep), FusionEngine.infer_sweep / infer_subset x 15, ece_bins; "
                           "predictions and bin statistics of every subset copied back"},
            "gpu_launches": None,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": tf / peaks["tflops"], "traffic": None, "peak_source": peaks["src"],
                         "kernel": "whole sweep; algorithmic FLOPs per window-evaluation: "
                                   + ("3 553 792 (dense path: every subset runs the full forward)" if args.no_mask_hint else
                                      "1 123 635 (mask-aware: only pairs with both modalities present, SURVEY 8d config 3)")},
            "accuracy_bins_total": total // (steps * repeats)}
    before = pkg.lib().msf_launch_count()
    sweep()
    torch.cuda.synchronize()
    line["gpu_launches"] = int(pkg.lib().msf_launch_count() - before) if args.no_graph else None
    line["gpu_launches_per_step"] = ("1 projection kernel + 15 x (folded pair GEMM + head kernel) + 15 ece_bin_kernel" if folded
                                     else "45 forward-pass kernels (3 per subset, CUDA graphs) + 15 ece_bin_kernel")
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _cpu_sweep_baseline(torch, subsets)
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def _cpu_sweep_baseline(torch, subsets, sample=8192):
    """The reference arithmetic of the sweep on the host: oracle forward (eval mode) + softmax/max + 15-bin ECE
    binning for every subset, on a bounded sample of windows."""
    from oracle import ece_oracle, fusion_oracle
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    fusion = importlib.import_module("fusion")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    names = list(DIMS)
    feats, _, labels = synthetic_batch(torch, 1234, sample)
    xs = dict(zip(names, feats))
    edges = ece_oracle.linspace_f32(15).astype("float64")

    def one_sweep():
        with torch.no_grad():
            for sub in subsets:
                mask = torch.zeros(sample, len(names))
                mask[:, list(sub)] = 1.0
                zeroed = {m: (x if i in sub else torch.zeros_like(x)) for i, (m, x) in enumerate(xs.items())}  # eval.py:401-404
                logits, _ = fusion_oracle.hybrid_fusion_forward(sd, names, HEADS, zeroed, mask)
                conf, pred = fusion_oracle.softmax_conf_pred(logits)
                ece_oracle.bin_masks(conf.numpy(), pred.numpy(), labels.numpy(), edges)

    one_sweep()
    times = []
    t_end = time.perf_counter() + 12.0
    while time.perf_counter() < t_end and len(times) < 20:
        t0 = time.perf_counter()
        one_sweep()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": len(subsets) * sample / med, "unit": "windows/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} sweeps of 15 subsets x {sample} windows (oracle fp32 forward + softmax/max + numpy "
                      f"binning), median; a reported baseline, not the target"}


def run_raw_infer(args):
    """Boundary E of SURVEY.md §8d: inference from RAW windows — 4 single-layer LSTM SequenceEncoders
    (T = 1024; F = 17, 17, 17, 1; hidden 256) -> Linear(256, 128) -> LayerNorm -> HybridFusion -> softmax/argmax,
    B = 4096 windows on one GPU.  The recurrence runs on msf_lstm_forward (ONE persistent tcgen05 launch over all
    time steps for the four encoders, lstm_seq.cu); the same pass with the library (cuDNN) recurrence is timed
    beside it."""
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    engine_mod = importlib.import_module(PKG + ".engine")
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    fusion = importlib.import_module("fusion")
    encoders = importlib.import_module("encoders")
    B, T, HID = BATCH, 1024, 256
    feats_in = {"imu_hand": 17, "imu_chest": 17, "imu_ankle": 17, "heart_rate": 1}
    torch.manual_seed(0)
    encs = {m: encoders.SequenceEncoder(f, hidden_dim=HID, output_dim=128, num_layers=1, encoder_type="lstm",
                                        dropout=DROPOUT).to(dev).eval() for m, f in feats_in.items()}
    norms = {m: torch.nn.LayerNorm(128).to(dev) for m in feats_in}
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT).eval()
    eng = engine_mod.FusionEngine(model, B, precision="bf16", use_graph=True)
    g = torch.Generator(device=dev).manual_seed(1234)
    xs = {m: torch.randn(B, T, f, device=dev, generator=g) for m, f in feats_in.items()}
    mask = torch.ones(B, len(feats_in), device=dev)
    packed_w = [ops.lstm_pack_weights(e.rnn.weight_ih_l0, e.rnn.weight_hh_l0, e.rnn.bias_ih_l0, e.rnn.bias_hh_l0)
                for e in encs.values()]

    def tail(hs):
        with torch.no_grad():
            enc_out = [norms[m](encoders._dense(encs[m].projection, h)) for m, h in zip(feats_in, hs)]
        return eng.infer(enc_out, mask)

    # the recurrence launch is captured once into a CUDA graph over static packed-input buffers
    static_x = [ops.lstm_pack_input(x) for x in xs.values()]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        ops.lstm_forward(static_x, packed_w, HID)                           # warm-up outside capture
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    lstm_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(lstm_graph, stream=side):
        static_h = ops.lstm_forward(static_x, packed_w, HID)

    def ours(src=None, after_pack=None):
        for dst, x in zip(static_x, src if src is not None else xs.values()):   # bf16, time-major, padded: part of the pass
            dst[:, :, :x.shape[2]] = x.transpose(0, 1)
        if after_pack is not None:
            after_pack()
        lstm_graph.replay()
        return tail(static_h)

    def library():
        with torch.no_grad():
            hs = [encoders._rnn_fp32(encs[m].rnn, xs[m])[1][0][-1] for m in feats_in]
        return tail(hs)

    def timed(fn, steps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps

    steps = max(1, min(args.steps, 5))
    logits_a = ours()[0].clone()
    logits_b = library()[0].clone()
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(ours, steps)
    clocks = sampler.stop()
    ms_lib = timed(library, steps)
    ms_lstm = timed(lstm_graph.replay, steps)   # the recurrence launch alone
    lstm_flop = B * sum(T * 2 * (f + HID) * 4 * HID for f in feats_in.values())
    lstm_only = {"ms": ms_lstm, "traffic": _lstm_traffic("lstm_seq_kernel_inference", T),
                 "tflops": lstm_flop / (ms_lstm * 1e-3) / 1e12,
                 "frac": lstm_flop / (ms_lstm * 1e-3) / 1e12 / _peaks()["tflops"],
                 "us_per_time_step": ms_lstm * 1e3 / T}
    # e2e: raw windows from pinned host memory, predictions and confidences read back, every pass
    host_x = {m: x.cpu().pin_memory() for m, x in xs.items()}
    host_pred = torch.zeros(B, dtype=torch.int64).pin_memory()
    host_conf = torch.zeros(B, dtype=torch.float32).pin_memory()

    pipe = HostPipeline(torch, list(host_x.values()), dev)

    def e2e():
        _, conf, pred = ours(pipe.next(), pipe.release)
        host_pred.copy_(pred, non_blocking=True)
        host_conf.copy_(conf, non_blocking=True)

    ms_e2e = timed(e2e, max(3, steps))
    del pipe
    # the same pass through the module-level drop-in (pipeline.EncodeFuse: SequenceEncoder modules grouped into one
    # launch of the recurrence, LayerNorm inside the projection kernel, HybridFusion module on the tensor-core path)
    import copy
    pipeline = importlib.import_module(PKG + ".pipeline")
    for e in encs.values():
        e.precision = "bf16"
    fus_mod = copy.deepcopy(model).to(dev).eval()
    fus_mod.precision = "bf16"
    glue = pipeline.EncodeFuse(encs, fus_mod, norms).eval()

    def module_pass():
        with torch.no_grad():
            return glue(xs, mask)

    mod_logits = module_pass().clone()
    ms_mod = timed(module_pass, steps)
    h2d = sum(t.numel() * 4 for t in host_x.values())
    flop = B * sum(T * 2 * (f + HID) * 4 * HID + 2 * HID * 128 for f in feats_in.values()) + B * FLOP_FWD
    peaks = _peaks()
    tf = flop / (ms * 1e-3) / 1e12
    line = {"metric": METRIC, "value": B / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1, "steps": steps, "warmup": 2,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "inference from raw windows (SURVEY 8d boundary E): 4 LSTM encoders T=1024 + "
                                   "projection + LayerNorm + HybridFusion + softmax", "batch": B, "seq_len": T,
                       "l2": "1.1 GB of raw windows per pass >> 126 MiB L2"},
            "clocks": clocks,
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": B * 12,
                    "api": "raw fp32 windows copied from pinned host memory (2-slot pipeline on a copy stream: the copy of "
                           "pass i+1 overlaps pass i), SequenceEncoder recurrence on ops.lstm_forward, FusionEngine.infer, "
                           "predictions + confidences copied back, every pass"},
            "gpu_launches": (1 if HID <= 256 and not os.environ.get("MSF_LSTM_STEPS") else T) + 8,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": tf / peaks["tflops"], "traffic": _lstm_traffic("lstm_seq_kernel_inference", T), "peak_source": peaks["src"],
                         "kernel": "whole pass: lstm_seq_kernel (one persistent launch over the 1024 steps of the 4 "
                                   "encoders; MSF_LSTM_STEPS=1: one tc_gemm_kernel launch per step) + fusion forward",
                         "lstm_kernel": lstm_only},
            "library_recurrence": {"ms_per_step": ms_lib, "value": B / (ms_lib * 1e-3),
                                   "what": "same pass with torch.nn.LSTM (cuDNN, fp32 no-TF32) for the recurrence"},
            "module_api": {"ms_per_step": ms_mod, "value": B / (ms_mod * 1e-3),
                           "what": "pipeline.EncodeFuse(SequenceEncoder x 4, LayerNorm, HybridFusion)(raw windows, mask) "
                                   "in eval mode, precision bf16: the drop-in modules a caller of train.py / eval.py uses",
                           "max_abs_logit_diff_vs_engine": float((mod_logits - logits_a).abs().max())},
            "max_abs_logit_diff_vs_library": float((logits_a - logits_b).abs().max())}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _cpu_raw_infer_baseline(torch, encoders, fusion, feats_in, T, HID)
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def _cpu_raw_infer_baseline(torch, encoders_mod, fusion, feats_in, T, HID, sample=64):
    """The reference arithmetic of the raw-window pass on the host cores: oracle LSTM recurrence (fp32) of the four
    encoders + projection + LayerNorm + oracle fusion forward + softmax/argmax on a bounded sample of windows."""
    from oracle import encoder_oracle, fusion_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    encs = {m: encoders_mod.SequenceEncoder(f, hidden_dim=HID, output_dim=128, num_layers=1, encoder_type="lstm",
                                            dropout=DROPOUT).eval() for m, f in feats_in.items()}
    norms = {m: torch.nn.LayerNorm(128) for m in feats_in}
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    xs = {m: torch.randn(sample, T, f, generator=g) for m, f in feats_in.items()}
    mask = torch.ones(sample, len(feats_in))

    def one_pass():
        with torch.no_grad():
            enc = {}
            for m in feats_in:
                esd = {k: v.detach() for k, v in encs[m].state_dict().items()}
                h = encoder_oracle.sequence_encoder_forward(esd, xs[m], 1, "lstm")
                enc[m] = norms[m](h)
            logits, _ = fusion_oracle.hybrid_fusion_forward(sd, list(DIMS), HEADS, enc, mask)
            return torch.softmax(logits, 1).max(1)

    one_pass()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 10.0:
        one_pass()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": sample / dt, "unit": "windows/s", "cores": cores, "kind": "port",
            "sample": f"{n} passes of {sample} raw windows (T = {T}): oracle LSTM recurrence + projection + LayerNorm + "
                      "fusion forward + softmax/argmax; a reported baseline, not the target"}


def run_raw_train(args):
    """Training pass of the four LSTM SequenceEncoders from RAW windows (the first "next" row of SURVEY.md §8f; the
    reference's nn.LSTM call under autograd, src/encoders.py:135-166): forward in training mode (msf_lstm_forward
    keeping h / gates / cell states) + backward (msf_lstm_backward: persistent backward kernel + weight-gradient
    GEMMs), B = 4096 windows, T = --seq-len steps, one GPU.  The library (cuDNN fp32) forward + backward of the same
    encoders is timed beside it; gradients of both are compared."""
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    sys.path.insert(0, os.path.join(ROOT, PKG, "src"))
    encoders = importlib.import_module("encoders")
    B, T, HID = BATCH, args.seq_len, 256
    feats_in = {"imu_hand": 17, "imu_chest": 17, "imu_ankle": 17, "heart_rate": 1}
    torch.manual_seed(0)
    encs = {m: encoders.SequenceEncoder(f, hidden_dim=HID, output_dim=128, num_layers=1, encoder_type="lstm",
                                        dropout=0.0).to(dev).train() for m, f in feats_in.items()}
    g = torch.Generator(device=dev).manual_seed(1234)
    xs = [torch.randn(B, T, f, device=dev, generator=g) for f in feats_in.values()]
    d_h = [torch.randn(B, HID, device=dev, generator=g) / B for _ in feats_in]
    weights = [(e.rnn.weight_ih_l0, e.rnn.weight_hh_l0, e.rnn.bias_ih_l0, e.rnn.bias_hh_l0) for e in encs.values()]

    def ours(src=None, after_pack=None):
        tapes = ops.lstm_train_forward(src if src is not None else xs, weights, HID)
        if after_pack is not None:
            after_pack()   # the raw windows are consumed (packed into the recurrence's operand layout)
        return ops.lstm_backward(tapes, d_h)

    def library():
        grads = []
        for e, x, d in zip(encs.values(), xs, d_h):
            for p in e.rnn.parameters():
                p.grad = None
            h = encoders._rnn_fp32(e.rnn, x)[1][0][-1]
            (h * d).sum().backward()
            grads.append((e.rnn.weight_ih_l0.grad, e.rnn.weight_hh_l0.grad, e.rnn.bias_ih_l0.grad))
        return grads

    def timed(fn, steps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps

    steps = max(1, min(args.steps, 5))
    ga = ours()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(ours, steps)
    clocks = sampler.stop()
    # where the pass spends its time: forward launch, backward recurrence, weight-gradient GEMMs
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    tapes = ops.lstm_train_forward(xs, weights, HID)
    ev[1].record()
    ops.lstm_backward(tapes, d_h)
    ev[2].record()
    torch.cuda.synchronize()
    ms_fwd, ms_bwd = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    del tapes
    # per-launch durations (msf_prof_*: CUDA events around every tensor-core launch of one pass)
    import ctypes
    nat = importlib.import_module(PKG + "._native")
    nat.check(nat.lib().msf_prof_enable(1))
    ours()
    buf = ctypes.create_string_buffer(1 << 16)
    nat.check(nat.lib().msf_prof_report(buf, len(buf)))
    nat.check(nat.lib().msf_prof_enable(0))
    launches = []
    for row in buf.value.decode().splitlines():
        label, cnt, msr, fl = row.split("\t")
        launches.append({"launch": label, "launches": int(cnt), "ms": float(msr),
                         "tflops": float(fl) / (float(msr) * 1e-3) / 1e12 if float(msr) > 0 else None})
    rel = None
    ms_lib = None
    try:
        torch.cuda.empty_cache()
        gb = library()
        rel = max(float((a[k] - b[k]).norm() / b[k].norm()) for a, b in zip(ga, gb) for k in range(3))
        ms_lib = timed(library, max(1, min(steps, 2)))
    except torch.cuda.OutOfMemoryError:
        pass
    fwd_flop = B * sum(T * 2 * (f + HID) * 4 * HID for f in feats_in.values())
    bwd_flop = B * sum(T * (2 * 4 * HID * HID + 2 * 4 * HID * (HID + 64)) for f in feats_in.values())
    peaks = _peaks()
    tf = (fwd_flop + bwd_flop) / (ms * 1e-3) / 1e12
    host_x = [x.cpu().pin_memory() for x in xs]
    host_g = [[torch.empty_like(t, device="cpu").pin_memory() for t in trip] for trip in ga]

    pipe = HostPipeline(torch, host_x, dev)

    def e2e():
        grads = ours(pipe.next(), pipe.release)
        for trip, htrip in zip(grads, host_g):
            for t, ht in zip(trip, htrip):
                ht.copy_(t, non_blocking=True)

    ms_e2e = timed(e2e, max(3, steps))
    del pipe
    # the whole model from raw windows through the module-level drop-in, as train.py's training_step wires it
    # (train.py:233-291,302-324): encoders (grouped launches of the recurrence) -> projection -> LayerNorm ->
    # HybridFusion -> cross entropy -> backward to every parameter
    module_api = None
    try:
        torch.cuda.empty_cache()
        pipeline = importlib.import_module(PKG + ".pipeline")
        fusion = importlib.import_module("fusion")
        for e in encs.values():
            e.precision = "bf16"
        fus_mod = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=DROPOUT).to(dev)
        fus_mod.precision = "bf16"
        norms = {m: torch.nn.LayerNorm(128).to(dev) for m in feats_in}
        glue = pipeline.EncodeFuse(encs, fus_mod, norms).train()
        feats = dict(zip(feats_in, xs))
        mask = torch.ones(B, len(feats_in), device=dev)
        labels = torch.randint(0, CLASSES, (B,), device=dev, generator=g)

        def module_step():
            for prm in glue.parameters():
                prm.grad = None
            loss = torch.nn.functional.cross_entropy(glue(feats, mask), labels, label_smoothing=SMOOTHING)
            loss.backward()
            return loss

        ms_mod = timed(module_step, max(1, min(steps, 3)))
        module_api = {"ms_per_step": ms_mod, "value": B / (ms_mod * 1e-3),
                      "what": "pipeline.EncodeFuse(SequenceEncoder x 4, LayerNorm, HybridFusion) in training mode, precision "
                              "bf16: forward + cross entropy + backward to every parameter through the drop-in modules",
                      "final_loss": float(module_step().detach())}
    except torch.cuda.OutOfMemoryError:
        pass
    line = {"metric": METRIC, "value": B / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1, "steps": steps, "warmup": 2,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "training pass of the 4 LSTM encoders from raw windows (SURVEY 8f): forward with tape + "
                                   "backward incl. weight gradients", "batch": B, "seq_len": T, "hidden": HID,
                       "l2": f"{sum(x.numel() for x in xs) * 4 / 1e9:.2f} GB of raw windows and "
                             f"{B * T * HID * (8 + 4 + 2) * 4 / 1e9:.1f} GB of tape per pass >> 126 MiB L2"},
            "clocks": clocks,
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(t.numel() * 4 for t in host_x),
                    "d2h_bytes_per_step": sum(t.numel() * 4 for trip in host_g for t in trip),
                    "api": "raw fp32 windows copied from pinned host memory (2-slot pipeline on a copy stream: the copy of "
                           "pass i+1 overlaps pass i), ops.lstm_train_forward + ops.lstm_backward, all parameter gradients "
                           "copied back, every pass"},
            "gpu_launches": 2 + 3 * len(feats_in),
            "phases_ms": {"forward (lstm_seq_kernel<true>, incl. input / weight packing)": ms_fwd,
                          "backward (lstm_bwd_kernel + weight-gradient GEMMs + reduction)": ms_bwd},
            "per_launch": launches,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": tf / peaks["tflops"], "traffic": (_lstm_traffic("lstm_seq_kernel_training", T) or 0) + (_lstm_traffic("lstm_bwd_kernel", T) or 0) or None, "peak_source": peaks["src"],
                         "kernel": "whole pass: lstm_seq_kernel<true> + lstm_bwd_kernel + tc_gemm_kernel<true> (weight gradients)"},
            "library_recurrence": None if ms_lib is None else {
                "ms_per_step": ms_lib, "value": B / (ms_lib * 1e-3),
                "what": "torch.nn.LSTM forward + autograd backward (cuDNN, fp32 no-TF32) of the same 4 encoders"},
            "max_rel_gradient_diff_vs_library": rel, "module_api": module_api}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _cpu_raw_train_baseline(torch, encoders, feats_in, T, HID)
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def _cpu_raw_train_baseline(torch, encoders_mod, feats_in, T, HID, sample=32):
    """The reference arithmetic of the same pass on the host cores: oracle LSTM recurrence (fp32) of the four encoders
    forward + autograd backward on a bounded sample of windows."""
    from oracle import encoder_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    encs = {m: encoders_mod.SequenceEncoder(f, hidden_dim=HID, output_dim=128, num_layers=1, encoder_type="lstm",
                                            dropout=0.0).train() for m, f in feats_in.items()}
    g = torch.Generator().manual_seed(7)
    xs = {m: torch.randn(sample, T, f, generator=g) for m, f in feats_in.items()}

    def one_pass():
        for m in feats_in:
            sd = {k: v.detach().clone().requires_grad_(True) for k, v in encs[m].state_dict().items()}
            encoder_oracle.lstm_last_hidden(sd, "rnn", xs[m], 1).sum().backward()

    one_pass()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 10.0:
        one_pass()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": sample / dt, "unit": "windows/s", "cores": cores, "kind": "port",
            "sample": f"{n} passes of {sample} raw windows (T = {T}): oracle LSTM recurrence of the 4 encoders, forward + "
                      "autograd backward; a reported baseline, not the target"}


def run_ece(args):
    """ECE / reliability binning kernel (uncertainty.py:84-171) on N = 2^28 samples resident in HBM:
    20 B/sample (f32 conf + i64 pred + i64 label), HBM-bound; buffers (5.4 GB) >> L2.  `e2e`: 2^25 samples from
    pinned host memory through ops.ece_bin with the statistics read back (PCIe-bound by construction);
    `cpu_baseline`: the C oracle (oracle/ece_oracle.c) on the host."""
    import numpy as np
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    n = 1 << 28
    g = torch.Generator(device=dev).manual_seed(1234)
    conf = torch.rand(n, device=dev, generator=g)
    pred = torch.randint(0, CLASSES, (n,), device=dev, generator=g)
    label = torch.randint(0, CLASSES, (n,), device=dev, generator=g)
    edges = torch.linspace(0, 1, 16).double().tolist()
    out = torch.zeros(3, 15, dtype=torch.int64, device=dev)
    steps = max(1, min(args.steps, 20))
    sampler = ClockSampler(0)
    sampler.start()
    out.zero_()
    ms, repeats = _timed_repeats(torch, lambda: ops.ece_bin(conf, pred, label, edges, out=out), steps, max(3, args.warmup))
    clocks = sampler.stop()
    assert int(out[0].sum().item()) == n * (steps * repeats + max(3, args.warmup)), "every in-range confidence lands in exactly one bin"
    # end to end: host buffers in, statistics out
    ne = 1 << 25
    h_conf, h_pred, h_label = (t[:ne].cpu().pin_memory() for t in (conf, pred, label))
    d_conf, d_pred, d_label = (torch.empty_like(t[:ne]) for t in (conf, pred, label))
    h_out = torch.zeros(3, 15, dtype=torch.int64).pin_memory()

    def e2e():
        d_conf.copy_(h_conf, non_blocking=True)
        d_pred.copy_(h_pred, non_blocking=True)
        d_label.copy_(h_label, non_blocking=True)
        out.zero_()
        ops.ece_bin(d_conf, d_pred, d_label, edges, out=out)
        h_out.copy_(out, non_blocking=True)

    ms_e2e, rep_e2e = _timed_repeats(torch, e2e, 3, 2, min_seconds=0.3)
    peaks = _peaks()
    gbs = 20.0 * n / (ms * 1e-3) / 1e9
    line = {"metric": "samples/sec ECE / reliability binning", "value": n / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic",
            "config": {"workload": "ECE binning, 15 bins, N = 2^28 samples"},
            "run": {"l2": "5.4 GB of inputs >> 126 MiB L2", "repeats": repeats, "timed_region_ms": ms * steps * repeats},
            "clocks": clocks,
            "e2e": {"value": ne / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "samples": ne,
                    "h2d_bytes_per_step": 20 * ne, "d2h_bytes_per_step": 3 * 15 * 8, "repeats": rep_e2e,
                    "api": "ops.ece_bin on device copies of pinned host arrays (3 H2D copies), statistics copied back"},
            "gpu_launches": steps, "gpu_launches_per_step": 1,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["src"],
                         "kernel": "ece_bin_kernel, 20 algorithmic bytes per sample"}}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            line["roofline"]["traffic"] = json.load(f).get("ece_bin_kernel_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    if not args.no_cpu_baseline:
        from oracle import ece_oracle
        ns = 1 << 24
        c_np, p_np, l_np = (t[:ns].cpu().numpy() for t in (conf, pred, label))
        e_np = np.asarray(edges, dtype=np.float64)
        cores = os.cpu_count() or 1
        ece_oracle.bin_masks_c(c_np, p_np, l_np, e_np, threads=cores)
        times = []
        t_end = time.perf_counter() + 10.0
        while time.perf_counter() < t_end and len(times) < 30:
            t0 = time.perf_counter()
            cnt, _, _ = ece_oracle.bin_masks_c(c_np, p_np, l_np, e_np, threads=cores)
            times.append(time.perf_counter() - t0)
        assert int(np.asarray(cnt).sum()) == ns
        med = statistics.median(times)
        line["cpu_baseline"] = {"value": ns / med, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{len(times)} passes over 2^24 samples, C oracle (oracle/ece_oracle.c: the "
                                          f"reference's per-bin mask passes), {cores} threads, median"}
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def profile_dominant_kernel(torch, pkg, eng, ring, ring_n, steps=12, chain_reps=16):
    """Per-launch durations measured live with CUDA events on the launching stream (msf_prof_*), eager launches
    over the same ring of batches; a device-side sleep in front of each step lets the host queue the whole step so
    the events bracket kernels, not launch gaps.  Two passes: (1) every tensor-core launch bracketed on its own
    (the event records serialise the step: no programmatic-dependent-launch overlap, so these are upper bounds);
    (2) the dominant kernel, the chained pair GEMMs, issued `chain_reps` times back to back inside ONE bracket
    (msf_prof_enable(R): the kernel only reads its operands and overwrites its outputs), so its average launch
    duration includes the prologue overlap it has in the real step and excludes the event records.  Operands are
    L2-resident in both, as in the step (they were written by the preceding kernel)."""
    import ctypes
    if eng.prec != pkg.native.MSF_PREC_BF16:
        return None
    lib = pkg.lib()
    snap = eng._snapshot()
    for i in range(3):
        eng.load_batch(*ring[i % ring_n])
        eng._enqueue_train_step()
    torch.cuda.synchronize()

    def one_pass(mode, n_steps):
        pkg.native.check(lib.msf_prof_enable(mode))
        for i in range(n_steps):
            eng.load_batch(*ring[(3 + i) % ring_n])
            torch.cuda._sleep(3_000_000)
            eng._enqueue_train_step()
        buf = ctypes.create_string_buffer(1 << 16)
        pkg.native.check(lib.msf_prof_report(buf, len(buf)))
        pkg.native.check(lib.msf_prof_enable(0))
        rows = []
        for line in buf.value.decode().splitlines():
            label, n, ms, flops = line.split("\t")
            rows.append({"launch": label, "launches": int(n), "us_per_launch": float(ms) * 1e3 / int(n),
                         "gflop_per_launch": float(flops) / int(n) / 1e9})
        return rows

    rows = one_pass(1, steps)
    chain_rows = [r for r in one_pass(chain_reps, 4) if "chain" in r["launch"]]
    eng._restore(snap)
    return {"steps": steps, "rows": rows, "chain_rows": chain_rows, "chain_reps": chain_reps}


def roofline_entry(kern, peaks, step_tflops):
    """roofline of the dominant kernel, chain_kernel (value_proj -> out_proj chain forward, its mirror
    backward: 2 launches, ~60 % of the step's FLOPs): algorithmic FLOPs of its launches / their summed
    duration from CUDA events.  `all_gemm_kernels` repeats it over every tensor-core launch of the step."""
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("chain_kernel_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001 - optional ncu-derived figure
        pass
    if not kern or not kern["rows"]:
        return {"bound": "tensor", "achieved": step_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": step_tflops / peaks["tflops"], "traffic": traffic, "peak_source": peaks["src"],
                "kernel": "whole fused train step (per-kernel events unavailable)"}

    def agg(rows):
        us = sum(r["us_per_launch"] * r["launches"] for r in rows)
        gf = sum(r["gflop_per_launch"] * r["launches"] for r in rows)
        n = sum(r["launches"] for r in rows)
        return us, gf, n

    chain = kern.get("chain_rows") or [r for r in kern["rows"] if "chain" in r["launch"]] or kern["rows"]
    us, gf, n = agg(chain)
    us_all, gf_all, n_all = agg(kern["rows"])
    achieved = gf * 1e9 / (us * 1e-6) / 1e12
    all_tf = gf_all * 1e9 / (us_all * 1e-6) / 1e12
    burst = peaks.get("tflops_burst")
    out = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
           "frac": achieved / peaks["tflops"], "traffic": traffic, "peak_source": peaks["src"],
           "kernel": ("chain2_kernel (tcgen05/TMEM/TMA chained pair GEMMs value_proj -> out_proj and their mirror "
                      "backward, half-pair software pipeline; 2 launches per step)") if SHAPE == "pamap2" else
                     "tc_gemm_kernel grouped launches of the pair projections (shape outside the chained kernel)",
           "avg_launch_us": us / n, "algorithmic_gflop_per_launch": gf / n,
           "how": "CUDA events around %d back-to-back launches of each of the step's two chain launches, inside an "
                  "eager train step (operands L2-resident as in the step)" % kern.get("chain_reps", 1),
           "all_gemm_kernels": {"launches_per_step": n_all // kern["steps"], "achieved": all_tf,
                                "frac": all_tf / peaks["tflops"], "us_per_step": us_all / kern["steps"],
                                "how": "each launch bracketed by its own event pair (serialised: upper bounds)"},
           "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / peaks["tflops"],
           "per_launch": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()} | {
               "launches": r["launches"] // kern["steps"]} for r in kern["rows"]]}
    if burst:
        out["frac_of_burst_peak"] = achieved / burst
        out["whole_step_frac_of_burst_peak"] = step_tflops / burst
    return out


def eng_launches(lib, before, eng):
    """Kernels of ONE step: with a CUDA graph the host-side counter only moves during
    capture, so count one eager enqueue of the same step instead."""
    import torch
    if not eng.use_graph:
        return int(lib.msf_launch_count() - before)
    snap = eng._snapshot()
    a = lib.msf_launch_count()
    eng._enqueue_train_step()
    torch.cuda.synchronize()
    n = int(lib.msf_launch_count() - a)
    eng._restore(snap)
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MSF_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the strong-scaling sub-record (BASELINE configs[3]: global batch 32768 split N ways)")
    ap.add_argument("--host-dtype", choices=["bf16", "fp32"], default="bf16",
                    help="train workload, e2e leg: dtype of the host feature batches (bf16 halves the PCIe bytes; the "
                         "tensor-core path rounds the features to bf16 in its first kernel either way)")
    ap.add_argument("--packed-host", action="store_true",
                    help="train workload, e2e leg: host batches in FusionEngine.pinned_batch() buffers (one H2D "
                         "transfer per batch instead of six); not measured yet")
    ap.add_argument("--steps-per-graph", type=int, default=0,
                    help="train workload: steps captured into one graph launch (FusionEngine.train_slots); "
                         "0 = 8 on one GPU when --steps >= 16, else 1")
    ap.add_argument("--no-fold", action="store_true",
                    help="infer_sweep: per-subset infer_subset (chained pair kernels) instead of the folded sweep")
    ap.add_argument("--no-mask-hint", action="store_true",
                    help="infer_sweep: run every subset through the dense path (per-row mask, no skipping)")
    ap.add_argument("--shape", default="pamap2", choices=["pamap2", "scaled"],
                    help="train workload: pamap2 = BASELINE configs[1] (4 x 128 -> hidden 256), scaled = BASELINE "
                         "configs[4] (8 modalities x 256 -> hidden 512, 8 heads, 11 classes)")
    ap.add_argument("--seq-len", type=int, default=1024, help="raw_train: time steps per window")
    ap.add_argument("--workload", default="train", choices=["train", "infer_sweep", "ece", "raw_infer", "raw_train"],
                    help="train = the benchmark proper (BASELINE configs[1]); the other two print extra evidence "
                         "lines for configs[2] and the ECE binning kernel (single GPU)")
    args = ap.parse_args()
    if args.shape == "scaled":
        use_scaled_shape()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer_sweep":
        run_infer_sweep(args)
    elif args.workload == "ece":
        run_ece(args)
    elif args.workload == "raw_infer":
        run_raw_infer(args)
    elif args.workload == "raw_train":
        run_raw_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
