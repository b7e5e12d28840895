#!/usr/bin/env python
"""bench.py -- train images/sec of the B200 ViT/DeiT step (BASELINE.json metric) and its reference arm.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port)

Workload (config.workload): BASELINE.json configs[1] -- DeiT-tiny (distilled, 3x224x224 synthetic tiles,
random-init weights) full train step: forward + 0.5*CE(cls)+0.5*CE(dist) (lightning_modules.py:459-461) +
backward + clip_grad_norm(1.0) + AdamW(lr 1e-4, wd 1e-5; configs/vit_optimizer_params.json), batch 256 per GPU.
N > 1 is weak scaling: every rank runs the same per-GPU batch and gradients are all-reduced (bucketed NCCL).
`--model vit_base` selects configs[3] (ViT-B/16) instead.

One JSON line is printed by rank 0; see README/DESIGN.md for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train images/sec"
UNIT = "images/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# algorithmic FLOPs per image of one train step (fwd+dgrad+wgrad, no recompute)
TRAIN_GFLOP = {"deit_tiny": 7.5637, "vit_base": 105.3784}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference_run(model_name: str, steps: int, warmup: int, batch: int = 32):
    """The reference's own arithmetic (oracle port, oracle/vit_oracle.py) on the host cores: forward + loss +
    backward + clip(1.0) + AdamW, fp32, all threads.  Returns (images/s, seconds/step, cores)."""
    import torch
    from oracle import vit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.DEIT_TINY if model_name == "deit_tiny" else O.VIT_BASE
    params = O.seeded_state_dict(cfg, 42)
    x, y = O.seeded_batch(cfg, batch, 42)
    state = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = O.train_step(params, x, y, cfg)
        O.clip_and_adamw_step(params, grads, state, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores


def gpu_eager_reference_run(model_name: str, steps: int, warmup: int, batch: int, precision: str):
    """Context number, NOT the reference arm the driver scores (that is the CPU run above): the same oracle port of the
    reference's arithmetic executed by PyTorch eager (ATen / cuBLAS / cuDNN kernels) on cuda:0 at the full per-GPU batch --
    SURVEY.md section 8(d) config 2 "compare against PyTorch eager of the reference class on the same GPU".
    precision: fp32 (true fp32 matmuls), tf32, or bf16 (torch.autocast).  Returns (images/s, seconds/step)."""
    import torch
    from oracle import vit_oracle as O
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = precision == "tf32"
    cfg = O.DEIT_TINY if model_name == "deit_tiny" else O.VIT_BASE
    params = {k: v.to(dev) for k, v in O.seeded_state_dict(cfg, 42).items()}
    x, y = O.seeded_batch(cfg, batch, 42)
    x, y = x.to(dev), y.to(dev)
    state = {}

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
            _, _, grads = O.train_step(params, x, y, cfg)
        O.clip_and_adamw_step(params, grads, state, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / steps
    return batch / sec, sec


# stdout carries exactly ONE JSON line: the process's fd 1 is pointed at stderr for the whole run (library banners such as
# "NCCL version ..." are C-level printf to fd 1) and the result line is written to a duplicate of the original stdout.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        ips, sec = gpu_eager_reference_run(args.model, args.steps, max(3, args.warmup), args.batch, args.ref_precision)
        emit({"impl": "reference", "reference_device": "cuda (PyTorch eager, oracle port)", "precision": args.ref_precision,
              "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": sec * 1e3, "higher_is_better": True, "data": "synthetic",
              "config": workload_config(args, cpu=True)})
        return
    batch = 32 if args.model == "deit_tiny" else 8
    ips, sec, cores = cpu_reference_run(args.model, args.steps, max(1, args.warmup), batch)
    sample = f"{args.steps} steps of batch {batch} (of the per-GPU batch {args.batch}), fwd+bwd+clip+AdamW fp32"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, cpu: bool = False):
    name = {"deit_tiny": "DeiT-tiny distilled (D=192, L=12, H=3, N=198)", "vit_base": "ViT-B/16 (D=768, L=12, H=12, N=197)"}[args.model]
    return {"workload": f"{name} full train step (fwd + 0.5CE+0.5CE | CE + bwd + clip 1.0 + AdamW), 3x224x224 synthetic tiles, "
                        f"batch {args.batch}/GPU, random-init weights",
            "model": args.model, "per_gpu_batch": args.batch, "global_batch": args.batch * max(1, args.gpus), "image": "3x224x224",
            "parallelism": f"dp{max(1, args.gpus)}", "mode": args.mode,
            "drop_rate": args.drop_rate, "drop_path_rate": args.drop_path_rate,
            "l2": "no flush needed: the step's working set (activations > 3 GB) is far larger than the 126 MB L2"}


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, profiles/): label of the
# live per-kernel profile -> case name of tools/kbench.py the capture was taken with (same shapes, batch 256)
NCU_CASES = {
    "deit_tiny": {
        "attention_bwd": "attn_bwd", "attention_fwd": "attn_fwd", "layernorm_bwd": "ln_bwd", "layernorm_fwd": "ln_fwd",
        "gemm_tcgen05[fwd] 50688x768x192 epi1": "gemm_gelu", "gemm_tcgen05[fwd] 50688x576x192 epi0": "gemm_qkv",
        "gemm_tcgen05[fwd] 50688x192x768 epi0": "gemm_fc2", "gemm_tcgen05[dgrad] 50688x192x768 epi0": "dgrad_fc1",
        "gemm_tcgen05[wgrad] 768x192x50688 epi3": "wgrad_fc1",
    },
    "vit_base": {
        "attention_bwd": "attn_bwd", "attention_fwd": "attn_fwd", "layernorm_bwd": "ln_bwd", "layernorm_fwd": "ln_fwd",
        "gemm_tcgen05[fwd] 50432x3072x768 epi1": "gemm_gelu", "gemm_tcgen05[fwd] 50432x2304x768 epi0": "gemm_qkv",
        "gemm_tcgen05[fwd] 50432x768x3072 epi0": "gemm_fc2", "gemm_tcgen05[dgrad] 50432x768x3072 epi0": "dgrad_fc1",
        "gemm_tcgen05[wgrad] 3072x768x50432 epi3": "wgrad_fc1",
    },
}


def ncu_traffic(model: str, label: str):
    f = ROOT / "profiles" / f"r01_ncu_traffic_{model}.json"
    case = NCU_CASES.get(model, {}).get(label)
    if case is None or not f.exists():
        return None, None
    ent = json.loads(f.read_text()).get(case)
    return (ent["dram_bytes"], f"profiles/{f.name}:{case}") if ent else (None, None)


# ------------------------------------------------------------------------------------------- per-kernel live profile
class KernelProfile:
    """Brackets every libvitk launch group of an EAGER step with CUDA events on the launching stream and aggregates
    device time, algorithmic FLOPs and algorithmic bytes per kernel class."""

    def __init__(self, ops_mod, by_shape: bool = False):
        import torch
        self.torch, self.ops, self.records, self._orig, self.by_shape = torch, ops_mod, [], {}, by_shape

    def _wrap(self, name, meta_fn):
        orig = getattr(self.ops, name)
        self._orig[name] = orig
        torch = self.torch

        def wrapped(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **k)
            e.record()
            label, flops, nbytes = meta_fn(*a, **k)
            self.records.append((label, flops, nbytes, s, e))
            return r
        setattr(self.ops, name, wrapped)

    def __enter__(self):
        es = lambda t: t.element_size()

        def gemm_meta(A, B, M, N, K, **k):
            a_mn, b_mn = k.get("a_mn", False), k.get("b_mn", False)
            kind = "wgrad" if (a_mn and b_mn) else ("dgrad" if b_mn else "fwd")
            out = k["out"]
            nbytes = 2 * (M * K + N * K) + M * N * es(out)
            if k.get("residual") is not None:
                nbytes += 4 * M * N
            if k.get("out2") is not None:
                nbytes += 2 * M * N
            if k.get("aux") is not None:
                nbytes += 2 * M * N
            if self.by_shape:
                return f"gemm_tcgen05[{kind}] {M}x{N}x{K} epi{k.get('epilogue', 0)}", 2.0 * M * N * K, float(nbytes)
            return f"gemm_tcgen05[{kind}]", 2.0 * M * N * K, float(nbytes)

        def attn_f(qkv, B, N, H, scale, **k):
            return "attention_fwd", 4.0 * B * H * N * N * 64, float(B * N * H * 64 * 2 * 4 + B * H * N * 4)

        def attn_b(qkv, out, dout, lse, B, N, H, scale, **k):
            return "attention_bwd", 8.0 * B * H * N * N * 64, float(B * N * H * 64 * 2 * 8 + B * H * N * 8)

        def ln_f(x, *a, **k):
            return "layernorm_fwd", 0.0, float(x.numel() * 6)

        def ln_b(dy, x, *a, **k):
            return "layernorm_bwd", 0.0, float(x.numel() * 16)

        def colsum(x, out, **k):
            return "colsum16", 0.0, float(x.numel() * 2)

        def generic(label, nbytes_fn):
            return lambda *a, **k: (label, 0.0, float(nbytes_fn(*a, **k)))

        self._wrap("gemm", gemm_meta)
        self._wrap("attention_fwd", attn_f)
        self._wrap("attention_bwd", attn_b)
        self._wrap("layernorm_fwd", ln_f)
        self._wrap("layernorm_bwd", ln_b)
        self._wrap("colsum16", colsum)
        self._wrap("patchify", generic("patchify", lambda images, P, **k: images.numel() * 6))
        self._wrap("tokens_bwd", generic("tokens_bwd", lambda dx, *a, **k: dx.numel() * 6))
        self._wrap("head_fwd", generic("head_fwd", lambda x, *a, **k: 0))
        self._wrap("head_bwd", generic("head_bwd", lambda *a, **k: a[8].numel() * 6))
        self._wrap("loss_fwd_bwd", generic("loss", lambda *a, **k: 0))
        self._wrap("prefix_tokens_fwd", generic("prefix_tokens", lambda *a, **k: 0))
        self._wrap("grad_sqnorm", generic("grad_sqnorm", lambda g, s: g.numel() * 4))
        self._wrap("adamw_step", generic("adamw", lambda p, *a, **k: p.numel() * 30))
        return self

    def __exit__(self, *a):
        for n, f in self._orig.items():
            setattr(self.ops, n, f)

    def table(self):
        self.torch.cuda.synchronize()
        agg = {}
        for label, flops, nbytes, s, e in self.records:
            d = agg.setdefault(label, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += flops
            d["bytes"] += nbytes
        return agg


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import thyroid_vit_cnn_comparison_b200  # noqa: F401
    from thyroid_vit_cnn_comparison_b200 import ops, optim, parallel, training, vit

    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.manual_seed(42)
    if args.model == "deit_tiny":
        model = vit.create_deit_tiny(img_size=224, patch_size=16, in_chans=3, num_classes=2, distilled=True,
                                     drop_rate=args.drop_rate, drop_path_rate=args.drop_path_rate)
    else:
        model = vit.create_vit_base(img_size=224, patch_size=16, in_chans=3, num_classes=2, drop_rate=args.drop_rate,
                                    drop_path_rate=args.drop_path_rate)
    model = model.to(dev).train()
    opt = optim.FusedAdamW(model, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
    teacher = None
    if args.mode == "distill":
        import torchvision
        torch.backends.cudnn.benchmark = True   # the teacher is ~170 small convolutions: let cuDNN pick per-shape algorithms
        teacher = torchvision.models.densenet169(weights=None, num_classes=2).to(dev).eval().to(torch.bfloat16)
        teacher = teacher.to(memory_format=torch.channels_last)
        for p in teacher.parameters():
            p.requires_grad = False
    reducer = parallel.BucketedAllReduce(bucket_mb=args.bucket_mb) if world > 1 else None
    use_graph = not args.no_graph and (world == 1 or not args.dp_eager)
    step = training.TrainStep(model, opt, args.batch, mode=args.mode, teacher=teacher, reducer=reducer, use_graph=use_graph)

    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host_imgs = [torch.rand(args.batch, 1, 224, 224, generator=g).expand(-1, 3, -1, -1).contiguous().pin_memory() for _ in range(n_host)]
    host_lbls = [torch.randint(0, 2, (args.batch,), generator=g).pin_memory() for _ in range(n_host)]
    h2d_bytes = host_imgs[0].numel() * 4 + host_lbls[0].numel() * 8
    stats_host = torch.zeros(8, dtype=torch.float32).pin_memory()
    d2h_bytes = stats_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    for i in range(warm):
        step(host_imgs[i % n_host], host_lbls[i % n_host])
    barrier()

    # launches per step (counted in the library, during one eager replay of the same step)
    ops.reset_launch_count()
    saved_graph, step.use_graph = step.use_graph, False
    step.run()
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    step.use_graph = saved_graph

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    with ClockSampler(local) as clk:
        # (1) inputs resident in HBM
        step.load(host_imgs[0], host_lbls[0])
        ms_dev, _ = timed(lambda i: step.run(), args.steps)
        # (2) end to end through the public call: pinned-host batch in, step statistics out, every step
        def e2e_step(i):
            st = step(host_imgs[i % n_host], host_lbls[i % n_host])
            stats_host.copy_(st, non_blocking=True)
        ms_e2e, wall_e2e = timed(e2e_step, args.steps)
    clocks = clk.summary()
    loss_val = float(stats_host[0])

    images = args.batch * world * args.steps
    value = images / (ms_dev / 1e3)
    e2e_value = images / (ms_e2e / 1e3)

    # live per-kernel profile (eager, CUDA events on the launching stream)
    roof, kernels = None, None
    pk, pk_src = peaks()
    nprof = 3
    if rank != 0:
        # the eager profile steps issue gradient all-reduces: every rank has to take part in them
        saved_graph, step.use_graph = step.use_graph, False
        for _ in range(nprof):
            step.run()
        step.use_graph = saved_graph
        torch.cuda.synchronize()
    if rank == 0:
        saved_graph, step.use_graph = step.use_graph, False
        with KernelProfile(ops, by_shape=True) as prof:
            for _ in range(nprof):
                # ~30 ms of device-side spinning first: the host enqueues the whole eager step behind it, so every event
                # pair brackets device time only (no launch gaps of a host-bound eager step inside the per-kernel numbers)
                torch.cuda._sleep(60_000_000)
                step.run()
            tab_shape = prof.table()
        step.use_graph = saved_graph
        # per-class table (GEMMs of all shapes folded into fwd / dgrad / wgrad) unless --by-shape asks for the detail
        tab = {}
        for label, v in tab_shape.items():
            key = label if args.by_shape else label.split(" ")[0]
            d = tab.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            for k2 in d:
                d[k2] += v[k2]
        total_ms = sum(v["ms"] for v in tab.values())
        kernels = {}
        ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        for label, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
            sec = v["ms"] / 1e3
            ent = {"launches_per_step": v["launches"] // nprof, "ms_per_step": v["ms"] / nprof, "share": v["ms"] / total_ms}
            if v["flops"] > 0:
                ent["tflops"] = v["flops"] / sec / 1e12
                ent["frac_tensor_peak"] = ent["tflops"] / pk["bf16_tflops_sustained"]
            ent["gbs"] = v["bytes"] / sec / 1e9
            ent["frac_hbm_peak"] = ent["gbs"] / pk["hbm_gbs"]
            ent["intensity_flop_per_byte"] = (v["flops"] / v["bytes"]) if v["bytes"] else None
            kernels[label] = ent
        # the roofline object describes the single dominant kernel (one shape), not a class of launches
        top_label, tv_ = max(tab_shape.items(), key=lambda kv: kv[1]["ms"])
        sec_ = tv_["ms"] / 1e3
        top = {"share": tv_["ms"] / total_ms, "gbs": tv_["bytes"] / sec_ / 1e9,
               "tflops": tv_["flops"] / sec_ / 1e12 if tv_["flops"] else 0.0,
               "intensity_flop_per_byte": (tv_["flops"] / tv_["bytes"]) if tv_["bytes"] else None}
        if top.get("intensity_flop_per_byte") and top["intensity_flop_per_byte"] > ridge:
            roof = {"kernel": top_label, "bound": "tensor", "achieved": top["tflops"], "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": top["tflops"] / pk["bf16_tflops_sustained"], "traffic": None}
        else:
            roof = {"kernel": top_label, "bound": "hbm", "achieved": top["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": top["gbs"] / pk["hbm_gbs"], "traffic": None}
        roof["traffic"], roof["traffic_source"] = ncu_traffic(args.model, top_label)
        roof["peak_source"] = f"{pk_src} (MEASURED_PEAKS.json sustained figures: the kernel is timed inside a long step)"
        roof["avg_launch_ms"] = tv_["ms"] / tv_["launches"]
        roof["algorithmic_per_launch"] = {"flops": tv_["flops"] / tv_["launches"], "bytes": tv_["bytes"] / tv_["launches"]}
        roof["share_of_step"] = top["share"]
        roof["whole_step_tensor_frac"] = value / world * TRAIN_GFLOP[args.model] * 1e9 / (pk["bf16_tflops_sustained"] * 1e12)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = 32 if args.model == "deit_tiny" else 8
        k = 6 if args.model == "deit_tiny" else 3
        ips, sec, cores = cpu_reference_run(args.model, k, 2, b)
        cpu_base = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{k} steps of batch {b} of the same workload (oracle/vit_oracle.py, fp32, fwd+bwd+clip+AdamW)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands / fp32 accumulate (fp32 master weights + residual stream, dynamic loss scale)",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "cuda_graph": bool(use_graph), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base, "kernels": kernels,
            "loss": loss_val, "loss_scale": float(step.eng.amp[0].item()), "skipped_steps": float(step.eng.amp[3].item()),
        }
        emit(line)
    if world > 1:
        # teardown must never hold the box: drop captured graphs (they pin NCCL work), then leave; a watchdog ends the
        # process if the communicator teardown itself blocks (the result line is already out)
        def _bail():
            time.sleep(20)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        step.graphs = [None, None]
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="deit_tiny", choices=["deit_tiny", "vit_base"])
    ap.add_argument("--mode", default="ce", choices=["ce", "distill"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dp-graph", action="store_true", help="(default) N > 1: the step, NCCL all-reduces included, is one CUDA graph")
    ap.add_argument("--dp-eager", action="store_true", help="N > 1: launch the step eagerly instead of replaying a captured graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--drop-rate", type=float, default=0.0, help="nn.Dropout rate (BASELINE.json's workload: 0)")
    ap.add_argument("--drop-path-rate", type=float, default=0.0, help="stochastic depth rate (BASELINE.json's workload: 0)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the scored reference arm; cuda = PyTorch-eager context number on the same GPU")
    ap.add_argument("--ref-precision", default="fp32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--by-shape", action="store_true", help="split the GEMM rows of the kernel profile by shape")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
