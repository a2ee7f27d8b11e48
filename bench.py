#!/usr/bin/env python
"""Headline benchmark of the FP4 Linear hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mistral7b|c1|llama70b] [--tp-mode peer|nccl]

Workload (BASELINE.json configs[2], the one the metric is quoted on): the Mistral-7B linear stack at
batch 1 - 32 layers x {q 4096x4096, k 1024x4096, v 1024x4096, o 4096x4096, gate 14336x4096,
up 14336x4096, down 4096x14336}, blocksize 64, fp32 absmax, bf16 activations, random-init packed weights
(synthetic; there is no network for checkpoints).  One STEP = one decode token = 224 fused dequant-GEMVs
chained through their outputs (h -> q,k,v; q -> o; o -> gate,up; up -> down -> next layer), i.e. the
data dependence of a real decoder with the attention / norm / activation kernels (not part of this
library) left out.  The 3.93 GB of weights exceed L2 (126 MB) 31x, so every step streams from HBM.

metric / value: whole-job algorithmic GB/s = (0.5625 B per weight + activations) * steps / device time,
timed with CUDA events around K CUDA-graph replays.  `tok_per_s` = steps / time.
e2e: the same through the public API (TorchFP4Linear modules under GraphedCallable) with, every step,
the host->device copy of the input activation from pinned memory and the device->host read of the
result inside the timed region.
grouped_launches: the same stack with q/k/v and gate/up each issued as ONE grouped launch (TorchFP4LinearGroup /
fp4_b200_gemv_grouped; 128 launches per token) - an extension the reference does not have, reported beside
the headline, never instead of it.
N > 1 (torchrun): tensor parallel, Megatron style - q/k/v/gate/up column-parallel, o/down row-parallel - strong
scaling of the same workload.  --tp-mode peer (default): the row-parallel partial sums are pushed into every
rank's symmetric-memory buffer by the GEMV itself and summed while the next launch stages x (PeerExchange /
fp4_b200_gemv_grouped_tp): no collective launches except one all_reduce per token for the final hidden state;
the NCCL all_reduce-per-layer variant is timed in the same run and reported as `nccl_allreduce_variant`.
--workload llama70b: BASELINE config #4 shapes (80 layers, 38.5 GB of FP4 linears), for reference.

--impl reference: the UNMODIFIED reference CUDA extension (oracle/_ref, built from /root/reference/csrc)
driven eagerly exactly as its Python module drives it (gemv_fp4 per layer on the legacy stream; it
cannot be graph-captured).  If that build is absent the CPU oracle port is timed instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MISTRAL = dict(hidden=4096, inter=14336, kv=1024, layers=32)
LLAMA70B = dict(hidden=8192, inter=28672, kv=1024, layers=80)  # BASELINE config #4 shapes (38.5 GB of FP4 linears)
BLOCKSIZE = 64


_REAL_STDOUT = None


def emit(text: str) -> None:
    """The one JSON line, on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (text + "\n").encode())
    else:
        print(text, flush=True)


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


def layer_shapes(cfg):
    h, i, kv = cfg["hidden"], cfg["inter"], cfg["kv"]
    # (name, out_features N, in_features K)
    return [("q", h, h), ("k", kv, h), ("v", kv, h), ("o", h, h), ("gate", i, h), ("up", i, h), ("down", h, i)]


def gemv_bytes(N, K, batch=1, act=2):
    return N * K // 2 + 4 * (N * K // BLOCKSIZE) + batch * K * act + batch * N * act


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a thread every 4 ms (the timed
    region of the default run is ~80 ms, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0, enabled=True):
        self.rows, self.proc, self.index = [], None, index
        self.sm, self.mx, self.reasons, self.how = [], [], set(), None
        self._stop = threading.Event()
        self.enabled = enabled  # only the rank that prints samples: eight processes polling the driver perturb it

    def _nvml_loop(self, nv, h):
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001 - older binding name
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:  # noqa: BLE001
            self.how = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        self._stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        if self.how:
            self.thread.join(timeout=2)

    def summary(self):
        if self.how == "nvml":
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 4 ms period"}
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


def synth_layer(N, K, dev, gen):
    """Random packed nibbles + absmax scaled so that a chain of layers keeps O(1) activations."""
    packed = torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev, generator=gen)
    gain = 1.0 / (K ** 0.5 * 0.45)
    absmax = (torch.rand(N * K // BLOCKSIZE, device=dev, generator=gen) + 0.5) * gain
    return packed, absmax


def build_stack(cfg, dev, rank=0, tp=1, seed=0):
    """The model as TorchFP4Linear modules (public API); TP shards when tp > 1."""
    import torch_bnb_fp4
    from torch_bnb_fp4_b200 import bnb_compat, ext
    from torch_bnb_fp4_b200.parallel import ColumnParallelFP4Linear, RowParallelFP4Linear, shard_column, shard_row

    gen = torch.Generator(device=dev).manual_seed(seed)
    code = torch.tensor(ext.BNB_FP4_CODE, dtype=torch.float32, device=dev)
    layers, nbytes = [], 0
    for _ in range(cfg["layers"]):
        mods = {}
        for name, N, K in layer_shapes(cfg):
            nbytes += gemv_bytes(N, K)
            packed, absmax = synth_layer(N, K, dev, gen)  # identical on every rank (same seed)
            n, k = N, K
            if tp > 1:  # cut the shard out of the full buffers, then drop them
                if name in ("o", "down"):
                    packed, absmax, k = shard_row(packed, absmax, N, K, rank, tp, BLOCKSIZE)
                else:
                    packed, absmax, n = shard_column(packed, absmax, N, K, rank, tp, BLOCKSIZE)
            lin = bnb_compat.LinearFP4(k, n, bias=False)
            st = bnb_compat.QuantState(absmax=absmax, shape=(n, k), code=code, blocksize=BLOCKSIZE,
                                       quant_type="fp4", dtype=torch.bfloat16)
            lin.weight = bnb_compat.Params4bit(packed, requires_grad=False, quant_state=st,
                                               blocksize=BLOCKSIZE, compress_statistics=False, quant_type="fp4")
            mods[name] = torch_bnb_fp4.TorchFP4Linear(lin, name=name)
        layers.append(mods)
    return layers, nbytes


def make_step(layers, tp):
    import torch.distributed as dist

    def step(h):
        for m in layers:
            q = m["q"](h)
            m["k"](h)
            m["v"](h)
            o = m["o"](q)
            if tp > 1:
                dist.all_reduce(o)
            m["gate"](o)
            up = m["up"](o)
            h = m["down"](up)
            if tp > 1:
                dist.all_reduce(h)
        return h
    return step


def make_step_grouped(layers, tp):
    """Same computation with q/k/v and gate/up issued as ONE grouped launch each (4 launches per layer)."""
    import torch.distributed as dist

    import torch_bnb_fp4
    groups = [(torch_bnb_fp4.TorchFP4LinearGroup([m["q"], m["k"], m["v"]]),
               torch_bnb_fp4.TorchFP4LinearGroup([m["gate"], m["up"]]), m) for m in layers]

    def step(h):
        for qkv, gu, m in groups:
            q, _, _ = qkv(h)
            o = m["o"](q)
            if tp > 1:
                dist.all_reduce(o)
            _, up = gu(o)
            h = m["down"](up)
            if tp > 1:
                dist.all_reduce(h)
        return h
    return step


def make_step_peer(layers, exchange):
    """Tensor parallel without collective launches: row-parallel partial sums stay in peer (symmetric) memory and
    are summed by the next column-parallel launch while it stages x (grouped q/k/v and gate/up launches).
    A row-parallel layer whose local K is outside the streaming kernel (K/tp % 256 != 0)
    keeps the NCCL all_reduce, as does the model's final hidden state."""
    import torch.distributed as dist

    from torch_bnb_fp4_b200.parallel import fused_tp_group_forward as fwd

    peer_o = layers[0]["o"].quant_data.N % 256 == 0
    peer_down = layers[0]["down"].quant_data.N % 256 == 0

    def step(h):
        x = h
        last = len(layers) - 1
        for i, m in enumerate(layers):
            q, _, _ = fwd([m["q"], m["k"], m["v"]], x, exchange)
            if peer_o:
                o = fwd([m["o"]], q, exchange, produce=True)
            else:
                o = m["o"](q)
                dist.all_reduce(o)
            _, up = fwd([m["gate"], m["up"]], o, exchange)
            if peer_down and i < last:
                x = fwd([m["down"]], up, exchange, produce=True)
            else:
                x = m["down"](up)
                dist.all_reduce(x)
        return x
    return step


def time_events(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def cpu_baseline_sample(seconds=10.0):
    """Oracle port of dequantize_fp4 + matmul on the host cores: one 4096x4096 batch-1 layer, repeated."""
    import numpy as np

    import oracle

    N = K = 4096
    rng = np.random.default_rng(0)
    packed = rng.integers(0, 256, N * K // 2, dtype=np.uint8)
    absmax = (rng.random(N * K // BLOCKSIZE) * 0.1 + 0.01).astype(np.float32)
    x = rng.standard_normal((1, K)).astype(np.float32)
    code = oracle.bnb_code()
    oracle.linear_f32(x, packed, absmax, code, N, K, BLOCKSIZE)
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.linear_f32(x, packed, absmax, code, N, K, BLOCKSIZE)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or reps >= 2000:
            break
    return {"value": gemv_bytes(N, K) * reps / dt / 1e9, "unit": "GB/s", "cores": oracle.num_threads(),
            "kind": "port", "ms_per_op": dt / reps * 1e3,
            "sample": f"{reps} x one 4096x4096 blocksize-64 batch-1 layer (fp32 dequant + dot, OpenMP) in {dt:.1f} s"}


def traffic_from_profile():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "gemv_traffic.json")))["dram_bytes_per_launch_avg"]
    except Exception:  # noqa: BLE001
        return None


def run_ours(args, rank, world):
    import torch.distributed as dist

    from torch_bnb_fp4_b200.graph import GraphedCallable

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    cfg = dict(MISTRAL)
    if args.workload == "c1":
        cfg = dict(hidden=4096, inter=4096, kv=4096, layers=10)  # 70 x 4096x4096 layers = 660 MB > L2
    elif args.workload == "llama70b":
        cfg = dict(LLAMA70B)
    layers, nbytes = build_stack(cfg, dev, rank, world)
    launches_per_step = len(layers) * 7
    step = make_step(layers, world)
    h0 = torch.randn(1, cfg["hidden"], device=dev).bfloat16()
    tp_mode = "none" if world == 1 else "nccl all_reduce after every row-parallel layer"
    nccl_line = None
    if world > 1 and args.tp_mode == "peer":
        # headline for N > 1: the peer-memory exchange; the NCCL variant is timed beside it
        from torch_bnb_fp4_b200.parallel import PeerExchange
        runner_n = GraphedCallable(step, [h0], warmup=3)
        for _ in range(args.warmup):
            runner_n.graph.replay()
        dist.barrier(); torch.cuda.synchronize()
        t_n = time_events(runner_n.graph.replay, args.steps)
        v = torch.tensor([t_n], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        nccl_line = {"tok_per_s": args.steps / float(v.item()), "ms_per_step": float(v.item()) / args.steps * 1e3}
        ref_out = runner_n(h0).float().clone()
        del runner_n
        try:
            exchange = PeerExchange(cfg["hidden"], torch.bfloat16, dev)
            step_p = make_step_peer(layers, exchange)
            with torch.no_grad():
                step_p(h0)  # shapes outside the streaming kernel raise here (e.g. K/tp % 512 != 0 at tp 8)
            ok = torch.ones(1, device=dev)
        except Exception as e:  # noqa: BLE001
            print(f"[rank {rank}] peer-memory exchange unavailable ({e}); using NCCL", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() > 0:
            step = step_p
            tp_mode = ("row-parallel partial sums exchanged through peer (symmetric) memory and summed in the "
                       "consumer's x staging; q/k/v and gate/up grouped; one NCCL all_reduce per token for the "
                       "final hidden state")
        else:
            nccl_line = None
    runner = GraphedCallable(step, [h0], warmup=max(3, args.warmup))
    if nccl_line is not None:
        got = runner(h0).float()
        nccl_line["max_rel_diff_peer_vs_nccl"] = float((got - ref_out).abs().max() / ref_out.abs().max())
        exchange.check()
    host_in = torch.randn(1, cfg["hidden"]).bfloat16().pin_memory()
    host_out = torch.empty(1, cfg["hidden"], dtype=torch.bfloat16).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(t):
        if world > 1:
            v = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return float(v.item())
        return t

    for _ in range(args.warmup):
        runner.graph.replay()
    barrier()
    with ClockSampler(dev.index, enabled=(rank == 0)) as clk:
        t_dev = time_events(runner.graph.replay, args.steps)
    barrier()
    t_dev = maxrank(t_dev)

    # extension measured beside the headline: q/k/v and gate/up as grouped launches (128 launches per step)
    if world == 1:
        step_g = make_step_grouped(layers, world)
        runner_g = GraphedCallable(step_g, [h0], warmup=3)
        for _ in range(args.warmup):
            runner_g.graph.replay()
        barrier()
        t_grp = maxrank(time_events(runner_g.graph.replay, args.steps))
        # the same through group_projections(): the UNMODIFIED per-projection step function, q/k/v and gate/up
        # sharing launches behind the modules' backs
        import torch_bnb_fp4

        class _Block(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.q_proj, self.k_proj, self.v_proj, self.o_proj = m["q"], m["k"], m["v"], m["o"]
                self.gate_proj, self.up_proj, self.down_proj = m["gate"], m["up"], m["down"]

        blocks = [_Block(m) for m in layers]
        n_groups = sum(torch_bnb_fp4.group_projections(b) for b in blocks)

        def step_dropin(h):
            for b in blocks:
                q = b.q_proj(h)
                b.k_proj(h)
                b.v_proj(h)
                o = b.o_proj(q)
                b.gate_proj(o)
                up = b.up_proj(o)
                h = b.down_proj(up)
            return h

        runner_d = GraphedCallable(step_dropin, [h0], warmup=3)
        for _ in range(args.warmup):
            runner_d.graph.replay()
        t_drop = time_events(runner_d.graph.replay, args.steps)
        with torch.no_grad():
            for _ in range(2):
                step_dropin(h0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                step_dropin(h0)
            torch.cuda.synchronize()
            t_drop_eager = (time.perf_counter() - t0) / 10
        dropin = {"groups": n_groups, "tok_per_s": args.steps / t_drop, "eager_tok_per_s": 1.0 / t_drop_eager,
                  "what": "group_projections(block) on blocks whose forward calls q/k/v/o/gate/up/down one by one"}
    else:
        t_grp, dropin = None, None

    # end to end: pinned host input -> H2D -> replay -> D2H of the result, every step
    def e2e_step():
        out = runner(host_in)
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    t_e2e = maxrank(time.perf_counter() - t0)

    # eager public-API path (no graph), for reference
    with torch.no_grad():
        for _ in range(2):
            step(h0)
        barrier()
        t0 = time.perf_counter()
        n_eager = max(1, min(args.steps, 10))
        for _ in range(n_eager):
            hdev = h0.copy_(host_in, non_blocking=True)
            host_out.copy_(step(hdev), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        t_eager = maxrank(time.perf_counter() - t0) / n_eager
    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    gbs = nbytes * args.steps / t_dev / 1e9
    line = {
        "metric": {"mistral7b": "batch-1 FP4 GEMV HBM GB/s (Mistral-7B-shape decode linear stack)",
                   "llama70b": "batch-1 FP4 GEMV HBM GB/s (Llama-3-70B-shape decode linear stack)",
                   "c1": "batch-1 FP4 GEMV HBM GB/s (4096x4096 layers)"}[args.workload],
        "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "tok_per_s": args.steps / t_dev,
        "config": {"workload": f"{args.workload}: {len(layers)} layers x 7 bnb-FP4 linears, blocksize 64, fp32 absmax, "
                               "batch 1, bf16 activations, random-init packed weights",
                   "algorithmic_bytes_per_step": nbytes, "launches_per_step": launches_per_step,
                   "l2_policy": "inputs larger than L2 (weights stream once per step)",
                   "parallelism": f"tp{world}" if world > 1 else "single GPU",
                   "timing": "CUDA events around CUDA-graph replays"},
        "e2e": {"value": nbytes * args.steps / t_e2e / 1e9, "unit": "GB/s", "tok_per_s": args.steps / t_e2e,
                "h2d_bytes_per_step": host_in.numel() * 2, "d2h_bytes_per_step": host_out.numel() * 2,
                "mode": "GraphedCallable replay of TorchFP4Linear modules; pinned H2D + D2H + sync each step",
                "eager_tok_per_s": 1.0 / t_eager, "eager_value": nbytes / t_eager / 1e9},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": gbs / world, "peak": peak, "unit": "GB/s",
                     "frac": gbs / world / peak, "peak_kind": peak_kind, "frac_of_nominal_8TBs": gbs / world / 8000.0,
                     "kernel": "gemv_stream_kernel (fused dequant-GEMV, IMMA u8 x s8)",
                     "avg_launch_us": t_dev / args.steps / launches_per_step * 1e6,
                     "algorithmic_bytes_per_launch_avg": nbytes / launches_per_step / world,
                     "traffic": traffic_from_profile()},
        "clocks": clk.summary(),
        "tp_collective": tp_mode,
        "grouped_launches": None if t_grp is None else {"value": nbytes * args.steps / t_grp / 1e9, "unit": "GB/s", "tok_per_s": args.steps / t_grp,
                             "ms_per_step": t_grp / args.steps * 1e3, "launches_per_step": len(layers) * 4,
                             "frac_of_peak": nbytes * args.steps / t_grp / 1e9 / world / peak,
                             "what": "extension: q/k/v and gate/up each issued as one fp4_b200_gemv_grouped launch "
                                     "(TorchFP4LinearGroup); same arithmetic up to fp32 summation order, 4 instead of 7 launches per layer",
                             "dropin": dropin},
    }
    if nccl_line is not None:
        nccl_line["value"] = nbytes / (nccl_line["ms_per_step"] * 1e-3) / 1e9
        line["nccl_allreduce_variant"] = nccl_line
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args.cpu_seconds)
    emit(json.dumps(line))


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = dict(MISTRAL)
    nbytes = sum(gemv_bytes(N, K) for _, N, K in layer_shapes(cfg)) * cfg["layers"]
    ref = None
    if torch.cuda.is_available():
        try:
            from oracle.build_ref import load_module
            ref = load_module()
        except Exception:  # noqa: BLE001
            ref = None
    base = {"impl": "reference", "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "metric": "batch-1 FP4 GEMV HBM GB/s (Mistral-7B-shape decode linear stack)"}
    if ref is None:
        cb = cpu_baseline_sample(args.cpu_seconds)
        base.update({"value": cb["value"], "ms_per_step": nbytes / (cb["value"] * 1e9) * 1e3,
                     "tok_per_s": cb["value"] * 1e9 / nbytes, "cpu_baseline": cb,
                     "config": {"workload": "CPU oracle port (reference CUDA extension not available): " + cb["sample"]},
                     "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(json.dumps(base))
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    code = torch.tensor([0.0, 5.208333333e-03, 0.66666667, 1.0, 0.33333333, 0.5, 0.16666667, 0.25,
                         -0.0, -5.208333333e-03, -0.66666667, -1.0, -0.33333333, -0.5, -0.16666667, -0.25],
                        dtype=torch.float32, device=dev)
    layers = []
    for _ in range(cfg["layers"]):
        mods = {}
        for name, N, K in layer_shapes(cfg):
            packed, absmax = synth_layer(N, K, dev, gen)
            mods[name] = (packed.t(), absmax, [N, K])  # B = self.A.t() as the reference passes it
        layers.append(mods)
    st = ref.bfloat16

    def gemv(m, x):  # reference torch_bnb_fp4/__init__.py:471-492 -> gemv_fp4
        B, am, shape = m
        return ref.gemv_fp4(x, B, am, code, BLOCKSIZE, st, shape)

    def step(h):
        for m in layers:
            q = gemv(m["q"], h)
            gemv(m["k"], h)
            gemv(m["v"], h)
            o = gemv(m["o"], q)
            gemv(m["gate"], o)
            up = gemv(m["up"], o)
            h = gemv(m["down"], up)
        return h

    h0 = torch.randn(1, cfg["hidden"], device=dev).bfloat16()
    host_in = torch.randn(1, cfg["hidden"]).bfloat16().pin_memory()
    host_out = torch.empty(1, cfg["hidden"], dtype=torch.bfloat16).pin_memory()
    for _ in range(max(1, args.warmup)):
        step(h0)
    torch.cuda.synchronize()
    with ClockSampler(0) as clk:
        t_dev = time_events(lambda: step(h0), args.steps)

    def e2e_step():
        h = h0.copy_(host_in, non_blocking=True)
        host_out.copy_(step(h), non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    t_e2e = time.perf_counter() - t0
    base.update({"value": nbytes * args.steps / t_dev / 1e9, "ms_per_step": t_dev / args.steps * 1e3,
                 "tok_per_s": args.steps / t_dev, "gpu_launches": 224 * args.steps, "clocks": clk.summary(),
                 "config": {"workload": "mistral7b: 32 layers x 7 bnb-FP4 linears, batch 1, bf16; UNMODIFIED reference "
                                        "CUDA extension (oracle/_ref) called eagerly per layer, legacy stream",
                            "algorithmic_bytes_per_step": nbytes},
                 "cpu_baseline": {"kind": "reference", "cores": os.cpu_count(), "value": nbytes * args.steps / t_dev / 1e9,
                                  "unit": "GB/s", "sample": "the reference has no CPU implementation of this path; this "
                                  "arm runs its CUDA extension on the GPU (see DESIGN.md)"},
                 "e2e": {"value": nbytes * args.steps / t_e2e / 1e9, "unit": "GB/s", "tok_per_s": args.steps / t_e2e,
                         "h2d_bytes_per_step": host_in.numel() * 2, "d2h_bytes_per_step": host_out.numel() * 2}})
    emit(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mistral7b", choices=["mistral7b", "c1", "llama70b"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tp-mode", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer-memory exchange fused into the consumer launch (default) or NCCL all_reduce")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if world > 1:
        # NCCL / torch print banners on stdout; the contract is ONE JSON line there: park stdout on stderr
        # until the line is printed
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        # symmetric-memory handles keep peer mappings alive; tearing the process group down under them can
        # block, so leave through _exit once every rank has passed the barrier
        os._exit(0)


if __name__ == "__main__":
    main()
