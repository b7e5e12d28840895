#!/usr/bin/env python
"""bench.py -- train images/sec of the B200 ViT/DeiT step (BASELINE.json metric) and its reference arm.

    python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA path, one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path on the host cores

Default line (config.workload): BASELINE.json configs[1] -- DeiT-tiny (distilled, 3x224x224 synthetic tiles, random-init
weights) full train step: forward + 0.5*CE(cls)+0.5*CE(dist) (lightning_modules.py:459-461) + backward +
clip_grad_norm(1.0) + AdamW(lr 1e-4, wd 1e-5; configs/vit_optimizer_params.json), batch 256 per GPU; and, because the metric
names both models, the same measurement of ViT-B/16 (configs[3]) as `models.vit_base` (>= 100 steps, own clocks / roofline /
kernels).  `--model X` measures one model alone.  N > 1 is weak scaling: every rank runs the same per-GPU batch and gradients
are all-reduced (bucketed NCCL, overlapped with backward).
    --mode distill    configs[2]: frozen DenseNet169 teacher forward + DeiT-tiny student step with the fused KL/CE loss
    --mode ensemble   configs[4]: 5-fold ensemble inference + attention rollout, folds sharded over the ranks (strong scaling)
    --dtype bf16      the north star's literal operand format (DESIGN.md section 3 explains why fp16 is the default)

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
REF_BATCH = {"deit_tiny": 32, "vit_base": 8}     # bounded CPU sample of the per-GPU batch


def _reference_classes(model_name: str):
    """The reference's OWN classes (unmodified, loaded from /root/reference by oracle/ref_loader.py) when that tree is
    mounted -- i.e. in the authoring container; the GPU box has no /root/reference, there the oracle port runs."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        _, vitm, deit = ref_loader.load()
        if model_name == "deit_tiny":
            return lambda: deit.create_deit_tiny(img_size=224, patch_size=16, in_chans=3, num_classes=2, distilled=True, pretrained=False)
        return lambda: vitm.create_vit_base(img_size=224, patch_size=16, in_chans=3, num_classes=2, drop_path_rate=0.0)
    except Exception:
        return None


def cpu_reference_run(model_name: str, steps: int, warmup: int, batch: int, mode: str = "ce"):
    """The reference's arithmetic on the host cores, fp32, all threads.  Returns (images/s, seconds/step, cores, kind).
    mode 'ce' / 'distill': forward + loss + backward + clip(1.0) + AdamW (distill: + a frozen torchvision DenseNet169 forward);
    mode 'ensemble': 5 eval forwards + softmax mix + attention rollout of each member.
    kind 'reference' = the reference's own classes drove the step; 'port' = oracle/vit_oracle.py."""
    import torch
    from oracle import vit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.DEIT_TINY if model_name == "deit_tiny" else O.VIT_BASE
    x, y = O.seeded_batch(cfg, batch, 42)
    times = []
    factory = _reference_classes(model_name) if mode == "ce" else None
    if factory is not None:
        torch.manual_seed(42)
        model = factory().train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            out = model(x)
            loss = O.classification_loss(out, y)              # lightning_modules.py:455-465
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        return batch / sec, sec, cores, "reference"
    params = O.seeded_state_dict(cfg, 42)
    if mode == "ensemble":
        members = [O.seeded_state_dict(cfg, 42 + f) for f in range(5)]
        w = torch.full((5,), 0.2)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                logits = []
                for sd in members:
                    maps = []
                    lg = O.forward(sd, x, cfg, training=False, attn_out=maps)
                    logits.append(lg)
                    O.attention_rollout(torch.stack(maps), "mean")
                O.ensemble_predict(torch.stack(logits), w)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        return batch / sec, sec, cores, "port"
    teacher = None
    if mode == "distill":
        import torchvision
        torch.manual_seed(0)
        teacher = torchvision.models.densenet169(weights=None, num_classes=2).eval()
    state = {}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tl = None
        if teacher is not None:
            with torch.no_grad():
                tl = teacher(x)
        _, _, grads = O.train_step(params, x, y, cfg, teacher_logits=tl)
        O.clip_and_adamw_step(params, grads, state, lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores, "port"


def gpu_eager_reference_run(model_name: str, steps: int, warmup: int, batch: int, precision: str):
    """Context number, NOT the reference arm the driver scores (that is the CPU run above): the same oracle port of the
    reference's arithmetic executed by PyTorch eager (ATen / cuBLAS / cuDNN kernels) on cuda:0 at the full per-GPU batch --
    SURVEY.md section 8(d) config 2 "compare against PyTorch eager of the reference class on the same GPU".
    precision: fp32 (true fp32 matmuls), tf32, or bf16 (torch.autocast).  Returns (images/s, seconds/step)."""
    import torch
    from oracle import vit_oracle as O
    dev = torch.device("cuda", torch.cuda.current_device())
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
    model = args.model or "deit_tiny"
    if args.ref_device == "cuda":
        ips, sec = gpu_eager_reference_run(model, args.steps, max(3, args.warmup), args.batch, args.ref_precision)
        emit({"impl": "reference", "reference_device": "cuda (PyTorch eager, oracle port)", "precision": args.ref_precision,
              "metric": metric_name(args), "value": ips, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": sec * 1e3, "higher_is_better": True, "data": "synthetic",
              "config": workload_config(args, model)})
        return
    batch = REF_BATCH[model]
    ips, sec, cores, kind = cpu_reference_run(model, args.steps, max(1, args.warmup), batch, args.mode)
    what = {"ce": "fwd+bwd+clip+AdamW", "distill": "frozen DenseNet169 fwd + student fwd+bwd+clip+AdamW",
            "ensemble": "5 eval forwards + softmax mix + 5 rollouts"}[args.mode]
    sample = (f"{args.steps} steps of batch {batch} (a bounded sample of the per-GPU batch {args.batch}), {what}, fp32, "
              + ("the reference's own classes (oracle/ref_loader.py)" if kind == "reference" else "oracle/vit_oracle.py port"))
    cfg = workload_config(args, model)
    cfg["per_gpu_batch_run_by_this_arm"] = batch          # what this CPU arm really ran; `per_gpu_batch` names the workload
    cfg["global_batch_run_by_this_arm"] = batch
    line = {
        "impl": "reference", "metric": metric_name(args), "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def metric_name(args) -> str:
    return "ensemble inference images/sec" if args.mode == "ensemble" else METRIC


def scaling_of(args) -> str:
    return "strong" if args.mode == "ensemble" else "weak"


MODEL_DESC = {"deit_tiny": "DeiT-tiny distilled (D=192, L=12, H=3, N=198)", "vit_base": "ViT-B/16 (D=768, L=12, H=12, N=197)"}


def workload_config(args, model: str):
    n = max(1, args.gpus)
    if args.mode == "ensemble":
        wl = (f"5-fold ensemble inference, {MODEL_DESC[model]} members (random init, seeds 42+f), eval forwards + "
              f"sum_f 0.2*softmax -> argmax + attention-rollout map [14,14] per fold, batch {args.batch} of 224x224 tiles, "
              f"fold f on rank f mod {n}")
        return {"workload": wl, "model": model, "folds": 5, "batch": args.batch, "image": "3x224x224 (single-channel uint16 tile, replicated on device)",
                "parallelism": f"fold-sharded over {n} rank(s)", "mode": "ensemble", "rollout": not args.no_rollout,
                "l2": "no flush needed: every member's activations + attention maps (> 1 GB) exceed the 126 MB L2"}
    loss = {"ce": "0.5CE+0.5CE | CE", "distill": "0.3*CE(cls,y) + 0.7*KL_T=3(dist || frozen DenseNet169 teacher)*9"}[args.mode]
    pre = "frozen torchvision DenseNet169 teacher fwd (bf16 channels_last, fused eval executor) + " if args.mode == "distill" else ""
    return {"workload": f"{MODEL_DESC[model]} full train step ({pre}fwd + {loss} + bwd + clip 1.0 + AdamW), 3x224x224 synthetic tiles "
                        f"(grayscale replicated to 3 channels), batch {args.batch}/GPU, random-init weights",
            "model": model, "per_gpu_batch": args.batch, "global_batch": args.batch * n, "image": "3x224x224",
            "input": ("single-channel uint16 tiles [B,224,224]; /65535 + replication to 3 channels fused into the patch-matrix kernel"
                      if args.input == "gray" else "fp32 [B,3,224,224]"),
            "parallelism": f"dp{n}", "mode": args.mode, "drop_rate": args.drop_rate, "drop_path_rate": args.drop_path_rate,
            "last_block": ("dense (VITK_DENSE_LAST_BLOCK=1)" if os.environ.get("VITK_DENSE_LAST_BLOCK", "0") == "1" or args.drop_rate > 0
                           or args.drop_path_rate > 0 else
                           "attn.proj / norm2 / Mlp of block L-1 run on the rows the classifier reads (cls, dist): same logits, loss and "
                           "gradients as the dense form, which VITK_DENSE_LAST_BLOCK=1 selects"),
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


def ncu_traffic_file(model: str):
    """Newest committed capture table for `model` (profiles/rNN_ncu_traffic_<model>.json)."""
    cands = sorted((ROOT / "profiles").glob(f"r*_ncu_traffic_{model}.json"))
    return cands[-1] if cands else None


def ncu_traffic(model: str, label: str):
    f = ncu_traffic_file(model)
    case = NCU_CASES.get(model, {}).get(label)
    if case is None or f is None:
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
        self._wrap("tiles_to_patches", generic("tiles_to_patches",
                                               lambda tiles, C, P, **k: tiles.numel() * (tiles.element_size() + 2 * C)))
        self._wrap("tokens_bwd", generic("tokens_bwd", lambda dx, *a, **k: dx.numel() * 6))
        self._wrap("gather_rows", generic("gather_rows", lambda src, n, out, **k: 2 * out.numel() * out.element_size()))
        self._wrap("expand_rows", generic("expand_rows", lambda src, n, out, **k: out.numel() * out.element_size()))
        self._wrap("head_fwd", generic("head_fwd", lambda x, *a, **k: 0))
        self._wrap("head_bwd", generic("head_bwd", lambda *a, **k: a[8].numel() * 6))
        self._wrap("loss_fwd_bwd", generic("loss", lambda *a, **k: 0))
        self._wrap("prefix_tokens_fwd", generic("prefix_tokens", lambda *a, **k: 0))
        self._wrap("grad_sqnorm", generic("grad_sqnorm", lambda g, *a, **k: g.numel() * 4))
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
def _kernel_tables(prof, nprof, model_name, pk, value_per_gpu, by_shape):
    """(kernels, roofline) from the live eager profile: per-class table, the dominant single-shape kernel as `roofline`,
    and `roofline.classes` = one entry per kernel class (share of the step, bound, achieved / peak, summed ncu traffic)."""
    tab_shape = prof.table()
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    total_ms = sum(v["ms"] for v in tab_shape.values())

    def entry(v):
        sec = v["ms"] / 1e3
        ent = {"launches_per_step": v["launches"] // nprof, "ms_per_step": v["ms"] / nprof, "share": v["ms"] / total_ms}
        if v["flops"] > 0:
            ent["tflops"] = v["flops"] / sec / 1e12
            ent["frac_tensor_peak"] = ent["tflops"] / pk["bf16_tflops_sustained"]
        ent["gbs"] = v["bytes"] / sec / 1e9
        ent["frac_hbm_peak"] = ent["gbs"] / pk["hbm_gbs"]
        ent["intensity_flop_per_byte"] = (v["flops"] / v["bytes"]) if v["bytes"] else None
        return ent

    def bound_of(ent):
        ai = ent.get("intensity_flop_per_byte")
        if ai and ai > ridge:
            return {"bound": "tensor", "achieved": ent["tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ent["tflops"] / pk["bf16_tflops_sustained"]}
        return {"bound": "hbm", "achieved": ent["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ent["gbs"] / pk["hbm_gbs"]}

    classes_tab, traffic_sum, traffic_ok = {}, {}, {}
    for label, v in tab_shape.items():
        key = label.split(" ")[0]
        d = classes_tab.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for k2 in d:
            d[k2] += v[k2]
        t, _src = ncu_traffic(model_name, label)
        if t is None:
            traffic_ok[key] = False
        else:
            traffic_sum[key] = traffic_sum.get(key, 0.0) + t * v["launches"] / nprof
            traffic_ok.setdefault(key, True)
    shown = tab_shape if by_shape else classes_tab
    kernels = {label: entry(v) for label, v in sorted(shown.items(), key=lambda kv: -kv[1]["ms"])}
    classes = []
    for key, v in sorted(classes_tab.items(), key=lambda kv: -kv[1]["ms"]):
        ent = entry(v)
        c = {"class": key, "launches_per_step": ent["launches_per_step"], "ms_per_step": ent["ms_per_step"], "share": ent["share"]}
        c.update(bound_of(ent))
        c["algorithmic_bytes_per_step"] = v["bytes"] / nprof
        c["traffic_per_step"] = traffic_sum.get(key) if traffic_ok.get(key) else None   # ncu dram bytes, all shapes of the class captured
        classes.append(c)
    top_label, tv_ = max(tab_shape.items(), key=lambda kv: kv[1]["ms"])
    top = entry(tv_)
    roof = {"kernel": top_label}
    roof.update(bound_of(top))
    roof["traffic"], roof["traffic_source"] = ncu_traffic(model_name, top_label)
    roof["peak_source"] = "MEASURED_PEAKS.json sustained figures (the kernel is timed inside a long step)" if pk.get("_src") == "measured" \
        else "fallback of /opt/skills/guides/B200_PROFILING.md"
    roof["avg_launch_ms"] = tv_["ms"] / tv_["launches"]
    roof["algorithmic_per_launch"] = {"flops": tv_["flops"] / tv_["launches"], "bytes": tv_["bytes"] / tv_["launches"]}
    roof["share_of_step"] = top["share"]
    roof["whole_step_tensor_frac"] = value_per_gpu * TRAIN_GFLOP[model_name] * 1e9 / (pk["bf16_tflops_sustained"] * 1e12)
    roof["whole_step_hbm_floor_ms"] = sum(v["bytes"] for v in tab_shape.values()) / nprof / (pk["hbm_gbs"] * 1e9) * 1e3
    roof["classes"] = classes
    return kernels, roof


def _timed(fn, k, world, dev):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push("timed")     # `ncu --nvtx --nvtx-include "timed/"` captures exactly the timed steps
    for i in range(k):
        fn(i)
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _host_tiles(args, rank, n_host):
    """Pinned host batches: single-channel tiles (uint16, as stored) or the replicated fp32 [B,3,H,W] batch."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    lbls = [torch.randint(0, 2, (args.batch,), generator=g).pin_memory() for _ in range(n_host)]
    if args.input == "gray":
        imgs = [torch.randint(0, 65536, (args.batch, 224, 224), generator=g, dtype=torch.int32).to(torch.uint16).pin_memory()
                for _ in range(n_host)]
    else:
        imgs = [torch.rand(args.batch, 1, 224, 224, generator=g).expand(-1, 3, -1, -1).contiguous().pin_memory() for _ in range(n_host)]
    return imgs, lbls


def measure_train(args, model_name, steps, rank, world, local, dev, primary):
    """One model's train-step record: value (inputs resident in HBM), e2e (pinned host tiles in, step statistics out, every
    step), clocks sampled during both, live per-kernel profile, roofline (+ classes), and -- N = 1, rank 0 -- the CPU baseline
    and the PyTorch-eager-on-this-GPU comparator."""
    import torch
    import torch.distributed as dist
    from thyroid_vit_cnn_comparison_b200 import ops, optim, parallel, training, vit

    torch.manual_seed(42)
    cdt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    if model_name == "deit_tiny":
        model = vit.create_deit_tiny(img_size=224, patch_size=16, in_chans=3, num_classes=2, distilled=True,
                                     drop_rate=args.drop_rate, drop_path_rate=args.drop_path_rate, compute_dtype=cdt)
    else:
        model = vit.create_vit_base(img_size=224, patch_size=16, in_chans=3, num_classes=2, drop_rate=args.drop_rate,
                                    drop_path_rate=args.drop_path_rate, compute_dtype=cdt)
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
    reducer = parallel.BucketedAllReduce(bucket_mb=args.bucket_mb, min_buckets=args.min_buckets) if world > 1 else None
    use_graph = not args.no_graph and (world == 1 or not args.dp_eager)
    step = training.TrainStep(model, opt, args.batch, mode=args.mode, teacher=teacher, reducer=reducer, use_graph=use_graph,
                              teacher_dtype=torch.bfloat16, input_format=args.input)

    n_host = 4
    host_imgs, host_lbls = _host_tiles(args, rank, n_host)
    h2d_bytes = host_imgs[0].numel() * host_imgs[0].element_size() + host_lbls[0].numel() * 8
    stats_host = torch.zeros(8, dtype=torch.float32).pin_memory()
    d2h_bytes = stats_host.numel() * 4

    warm = max(3, args.warmup)
    for i in range(warm):
        step(host_imgs[i % n_host], host_lbls[i % n_host])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # launches per step (counted in the library, during one eager replay of the same step)
    ops.reset_launch_count()
    saved_graph, step.use_graph = step.use_graph, False
    if reducer is not None:
        reducer.launched_log.clear()
    step.run()
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    step.use_graph = saved_graph
    dp = None
    if reducer is not None:
        log = list(reducer.launched_log)
        dp = {"buckets": len(reducer.buckets), "bucket_bytes": [4 * (e - s) for s, e in reducer.buckets],
              "payload": "fp32", "nccl_max_ctas": args.nccl_max_ctas or None, "launched_before_backward_ended": sum(1 for stage, _ in log if stage != "embed"),
              "launch_stages": [stage for stage, _ in log]}

    with ClockSampler(local) as clk:
        # (1) inputs resident in HBM
        step.load(host_imgs[0], host_lbls[0])
        ms_dev = _timed(lambda i: step.run(), steps, world, dev)

        # (2) end to end through the public call: pinned-host batch in, step statistics out, every step
        def e2e_step(i):
            st = step(host_imgs[i % n_host], host_lbls[i % n_host])
            stats_host.copy_(st, non_blocking=True)
        ms_e2e = _timed(e2e_step, steps, world, dev)
    clocks = clk.summary()
    loss_val = float(stats_host[0])

    images = args.batch * world * steps
    value = images / (ms_dev / 1e3)
    e2e_value = images / (ms_e2e / 1e3)

    # live per-kernel profile (eager, CUDA events on the launching stream)
    roof, kernels = None, None
    pk, pk_src = peaks()
    pk = dict(pk, _src=pk_src)
    nprof = 3
    if rank != 0:
        # the eager profile steps issue gradient all-reduces: every rank has to take part in them
        saved_graph, step.use_graph = step.use_graph, False
        for _ in range(nprof):
            step.run()
        step.use_graph = saved_graph
        torch.cuda.synchronize()
    else:
        saved_graph, step.use_graph = step.use_graph, False
        with KernelProfile(ops, by_shape=True) as prof:
            for _ in range(nprof):
                # ~30 ms of device-side spinning first: the host enqueues the whole eager step behind it, so every event
                # pair brackets device time only (no launch gaps of a host-bound eager step inside the per-kernel numbers)
                torch.cuda._sleep(60_000_000)
                step.run()
            kernels, roof = _kernel_tables(prof, nprof, model_name, pk, value / world, args.by_shape)
        step.use_graph = saved_graph

    loss_scale, skipped = float(step.eng.amp[0].item()), float(step.eng.amp[3].item())
    # release this model before the next one / the comparators run
    step.graphs = [None, None]
    del step, opt, model, teacher, reducer
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    cpu_base, eager = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if primary:
            b = REF_BATCH[model_name]
            k = 6 if model_name == "deit_tiny" else 3
            ips, sec, cores, kind = cpu_reference_run(model_name, k, 2, b, args.mode)
            cpu_base = {"value": ips, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{k} steps of batch {b} of the same workload ("
                                  + ("the reference's own classes" if kind == "reference" else "oracle/vit_oracle.py") + ", fp32)"}
        if args.mode == "ce":
            # comparator on THIS GPU: the reference's arithmetic (oracle port) run by PyTorch eager under torch.autocast(bf16)
            try:
                ips_e, sec_e = gpu_eager_reference_run(model_name, 3, 2, args.batch, "bf16")
                eager = {"value": ips_e, "unit": UNIT, "ms_per_step": sec_e * 1e3, "precision": "torch.autocast(bf16), fp32 master weights",
                         "what": "PyTorch eager (ATen/cuBLAS) run of the reference's arithmetic (oracle port) on this GPU, same batch, "
                                 "inputs resident, 3 timed steps -- context only, not the scored reference arm"}
            except Exception as e:  # noqa: BLE001 -- a comparator must never cost the measurement
                eager = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()

    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("fp16 operands / fp32 accumulate (fp32 master weights + residual stream, dynamic loss scale)" if args.dtype == "fp16"
                  else "bf16 operands / fp32 accumulate (fp32 master weights + residual stream)"),
        "data": "synthetic", "config": workload_config(args, model_name),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step),
        "cuda_graph": bool(use_graph), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
        "eager_same_gpu": eager, "dp": dp, "kernels": kernels,
        "loss": loss_val, "loss_scale": loss_scale, "skipped_steps": skipped,
    }
    return rec


def measure_ensemble(args, model_name, steps, rank, world, local, dev):
    """BASELINE.json config 5: F = 5 random-init members (seeds 42+f), eval mode, fold f on rank f mod world, uniform weights,
    one attention-rollout grid per fold.  A step classifies one batch with the whole ensemble; total work is fixed as N grows
    (strong scaling: at most 5 ranks hold a member)."""
    import torch
    import torch.distributed as dist
    from thyroid_vit_cnn_comparison_b200 import ensemble, ops, parallel, vit
    from thyroid_vit_cnn_comparison_b200.engine import GraySpec
    F = 5
    mine = parallel.shard_folds(F, rank, world)
    members = []
    for f in mine:
        torch.manual_seed(42 + f)
        if model_name == "deit_tiny":
            m = vit.create_deit_tiny(img_size=224, patch_size=16, in_chans=3, num_classes=2, distilled=True)
        else:
            m = vit.create_vit_base(img_size=224, patch_size=16, in_chans=3, num_classes=2, drop_path_rate=0.0)
        members.append(m.to(dev).eval())
    ens = ensemble.EnsembleInference(members, num_folds=F, rollout=not args.no_rollout)
    n_host = 4
    host_imgs, _ = _host_tiles(args, 0, n_host)            # every rank classifies the SAME batch (rank-0 seed)
    gray = GraySpec() if args.input == "gray" else None
    slots = [torch.empty_like(h, device=dev) for h in host_imgs[:2]]
    h2d_bytes = host_imgs[0].numel() * host_imgs[0].element_size()
    preds_host = torch.zeros(args.batch, dtype=torch.int64).pin_memory()
    grid_host = torch.zeros(F, args.batch, 14, 14, dtype=torch.float32).pin_memory() if not args.no_rollout else None
    d2h_bytes = preds_host.numel() * 8 + (grid_host.numel() * 4 if grid_host is not None else 0)

    def run(i, e2e):
        buf = slots[i & 1]
        if e2e:
            buf.copy_(host_imgs[i % n_host], non_blocking=True)
        out = ens(buf, gray=gray)
        if e2e:
            preds_host.copy_(out["preds"], non_blocking=True)
            if grid_host is not None:
                grid_host.copy_(out["rollout"], non_blocking=True)
        return out

    warm = max(3, args.warmup)
    for i in range(2):
        slots[i].copy_(host_imgs[i])
    for i in range(warm):
        run(i, True)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    run(0, False)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    with ClockSampler(local) as clk:
        ms_dev = _timed(lambda i: run(i, False), steps, world, dev)
        ms_e2e = _timed(lambda i: run(i, True), steps, world, dev)
    clocks = clk.summary()
    value = args.batch * steps / (ms_dev / 1e3)
    e2e_value = args.batch * steps / (ms_e2e / 1e3)
    pk, pk_src = peaks()
    # roofline of the dominant kernel of an eval forward with map emission: the attention kernel writes B*H*N*N fp32
    # probabilities per layer -- HBM-bound; achieved = algorithmic bytes of all member forwards' map writes / step time is a
    # LOWER bound of that kernel's rate (the step also contains the GEMMs), reported as such
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = 8
        ips, sec, cores, kind = cpu_reference_run(model_name, 2, 1, b, "ensemble")
        cpu_base = {"value": ips, "unit": UNIT, "cores": cores, "kind": kind,
                    "sample": f"2 steps of batch {b}: 5 eval forwards + softmax mix + 5 rollouts (oracle/vit_oracle.py, fp32)"}
    N, H, L = (198, 3, 12) if model_name == "deit_tiny" else (197, 12, 12)
    map_bytes = float(len(mine)) * L * args.batch * H * N * N * 4 * 2      # written by attention, read by the rollout fuse pass
    roof = {"kernel": "eval attention with fp32 map emission + rollout head-fusion read (per rank, all local members)",
            "bound": "hbm", "achieved": map_bytes / (ms_dev / steps / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": map_bytes / (ms_dev / steps / 1e3) / 1e9 / pk["hbm_gbs"], "traffic": None,
            "note": "whole-step lower bound: algorithmic attention-map bytes of this rank's members / step time"} if not args.no_rollout else None
    rec = {
        "metric": "ensemble inference images/sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp16 operands / fp32 accumulate", "data": "synthetic", "config": workload_config(args, model_name),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step), "cuda_graph": False,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base, "local_folds": mine,
    }
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import thyroid_vit_cnn_comparison_b200  # noqa: F401
    from thyroid_vit_cnn_comparison_b200 import parallel

    if args.nccl_max_ctas > 0:
        os.environ["NCCL_MAX_CTAS"] = str(args.nccl_max_ctas)
    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    primary = args.model or "deit_tiny"
    if args.mode == "ensemble":
        line = measure_ensemble(args, primary, args.steps, rank, world, local, dev)
    else:
        line = measure_train(args, primary, args.steps, rank, world, local, dev, primary=True)
        if args.model is None and args.mode == "ce" and not args.no_second_model:
            # BASELINE.json's metric names BOTH models: ViT-B/16 (configs[3]) rides in the same run as a sub-record,
            # at >= 100 steps (its step is power-state sensitive: a longer region averages that out)
            sub = measure_train(args, "vit_base", max(args.steps, args.second_model_steps), rank, world, local, dev, primary=False)
            line["models"] = {"vit_base": sub}
    if rank == 0:
        emit(line)
    if world > 1:
        # teardown must never hold the box: a watchdog ends the process if the communicator teardown itself blocks (the
        # result line is already out)
        def _bail():
            time.sleep(20)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
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
    ap.add_argument("--model", default=None, choices=["deit_tiny", "vit_base"],
                    help="default: DeiT-tiny as the line + ViT-B/16 as line['models']['vit_base'] (mode ce); naming a model measures it alone")
    ap.add_argument("--mode", default="ce", choices=["ce", "distill", "ensemble"],
                    help="ce: BASELINE configs 2/4; distill: config 3 (DenseNet169 teacher -> DeiT-tiny); ensemble: config 5")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (ensemble: the batch every rank classifies)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"], help="tensor-core operand format (DESIGN.md section 3)")
    ap.add_argument("--input", default="gray", choices=["gray", "nchw"],
                    help="gray: single-channel uint16 tiles, replicated on device; nchw: fp32 [B,3,H,W] batches")
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    ap.add_argument("--min-buckets", type=int, default=1,
                    help="N > 1: lower bound on the number of all-reduce buckets (a model smaller than --bucket-mb otherwise reduces "
                         "in ONE bucket after backward).  Measured on 8 B200 (profiles/r02_dp_scan_8gpu.txt): for DeiT-tiny's 22 MB "
                         "of gradients 1 bucket is fastest -- overlapped NCCL kernels take SMs from the persistent GEMMs")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="N > 1: cap the CTAs NCCL may use per collective (NCCL_MAX_CTAS) so that the overlapped all-reduces leave "
                         "the SMs to the persistent GEMM / attention kernels; 0 = NCCL's default")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dp-graph", action="store_true", help="(default) N > 1: the step, NCCL all-reduces included, is one CUDA graph")
    ap.add_argument("--dp-eager", action="store_true", help="N > 1: launch the step eagerly instead of replaying a captured graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-second-model", action="store_true", help="skip the ViT-B/16 sub-record of the default run")
    ap.add_argument("--second-model-steps", type=int, default=100)
    ap.add_argument("--no-rollout", action="store_true", help="ensemble mode: skip the attention-rollout maps")
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
