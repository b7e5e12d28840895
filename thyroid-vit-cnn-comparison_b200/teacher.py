"""Frozen-teacher fast path for the distillation step (reference:
`ThyroidDistillationModule.get_teacher_outputs`, lightning_modules.py:943-947 -- `self.teacher(images)` in eval mode under
no_grad, teacher = DenseNet169, src/models/cnn/densenet.py:24-45).

PyTorch eager spends 78 % of the DenseNet169 forward in memory-bound glue (measured on B200, batch 256, bf16 channels_last,
tools/teacher_profile.py: batch_norm 37 %, torch.cat / copies 29 %, relu 11 %; the cuDNN convolutions are 16 %).  A frozen
eval-mode network allows three exact rewrites:
  * the features of a dense block live in ONE preallocated NHWC buffer: a layer's 32 new channels are written next to the
    existing ones, so `torch.cat` (a copy of everything so far, every layer) disappears;
  * norm1 + relu1 over that concatenation is one pass of `vitk_affine_relu_nhwc` (eval BatchNorm is a per-channel affine map);
  * norm2 (and the stem's norm0) directly follow a convolution: they fold into its weights and bias, and the ReLU behind them
    rides on cuDNN's fused conv + bias + relu;
  * the stem's max pool and the transitions' average pools (`vitk_pool_nhwc`) store straight into the first channels of the next
    block's buffer.
The convolutions stay cuDNN calls (library GEMMs on a frozen network); parameters are snapshotted at construction, which is
what "frozen" means in the reference (`freeze_teacher`, lightning_modules.py:771-773).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _bn_affine(bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm as y = x * scale + shift (fp32)."""
    if not isinstance(bn, nn.BatchNorm2d) or bn.running_mean is None:
        raise TypeError("expected an nn.BatchNorm2d with running statistics")
    w = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(bn.running_mean, dtype=torch.float32)
    b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(bn.running_mean, dtype=torch.float32)
    scale = w / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = b - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """conv followed by eval BatchNorm -> (weight, bias) of the equivalent convolution (fp32)."""
    scale, shift = _bn_affine(bn)
    w = conv.weight.detach().float() * scale[:, None, None, None]
    b = shift if conv.bias is None else shift + conv.bias.detach().float() * scale
    return w, b


def is_supported(module: nn.Module) -> bool:
    """torchvision-style DenseNet: features.{conv0,norm0,pool0,denseblockN.denselayerM.{norm1,conv1,norm2,conv2},transitionN,norm5}
    + classifier."""
    f = getattr(module, "features", None)
    if f is None or not all(hasattr(f, n) for n in ("conv0", "norm0", "pool0", "denseblock1", "norm5")):
        return False
    if any(hasattr(f, n) for n in ("conv1", "norm1", "conv2", "norm2")):
        return False          # a deep-stem DenseNet (timm `stem_type='deep'`: conv0-norm0-conv1-norm1-conv2): not this executor's graph
    if not (isinstance(f.conv0, nn.Conv2d) and isinstance(f.norm0, nn.BatchNorm2d) and isinstance(f.pool0, nn.MaxPool2d)):
        return False
    return isinstance(getattr(module, "classifier", None), nn.Linear)


class FrozenDenseNet:
    """Callable replacement for `teacher(images)` of a frozen torchvision-style DenseNet (see module docstring)."""

    def __init__(self, module: nn.Module, dtype=torch.bfloat16, affine_relu: Optional[Callable] = None,
                 pool: Optional[Callable] = None, bottleneck: Optional[Callable] = None):
        if not is_supported(module):
            raise TypeError("FrozenDenseNet expects a torchvision-style DenseNet (features.denseblockN.denselayerM, classifier)")
        self.dtype = dtype
        self._affine_relu = affine_relu if affine_relu is not None else ops.affine_relu_nhwc
        self._pool = pool if pool is not None else ops.pool_nhwc
        # norm1 + relu1 + conv1 (1x1) + norm2 + relu2 of every dense layer as ONE tcgen05 GEMM over the block buffer
        # (ops.dense_bottleneck); with an injected affine_relu (the CPU restatement used by the tests) the two-step form runs
        self._bottleneck = bottleneck if bottleneck is not None else (ops.dense_bottleneck if affine_relu is None else None)
        f = module.features
        cl = lambda w: w.to(dtype).contiguous(memory_format=torch.channels_last)
        w0, b0 = _fold(f.conv0, f.norm0)
        # The 3-channel stem runs on a pre-tensor-core cuDNN kernel (1.8 ms at batch 256; zero-padding it to 8 channels was
        # measured: the library then picks a 2.8 ms kernel).  GEMM form: patch rows (vitk_im2col_rows, element order ky, kx, c)
        # against the filters reshaped the same way; the ReLU moves behind the max-pool (max and ReLU commute).
        self.stem = (cl(w0), b0.to(dtype), f.conv0.stride, f.conv0.padding)
        k0 = f.conv0.kernel_size
        self.stem_gemm = None
        self.stem_fused = True      # False: patch rows + vitk_gemm (the first GEMM form; kept for other stem shapes and for A/B runs)
        if k0[0] == k0[1] and f.conv0.stride[0] == f.conv0.stride[1] and f.conv0.padding[0] == f.conv0.padding[1] and w0.shape[0] % 8 == 0:
            wm = w0.permute(0, 2, 3, 1).reshape(w0.shape[0], -1)                       # [Cout, ky*kx*c]
            ld = (wm.shape[1] + 7) // 8 * 8
            wm = F.pad(wm, (0, ld - wm.shape[1])).to(dtype).contiguous()
            self.stem_gemm = {"w": wm, "b": b0.float().contiguous(), "k": int(k0[0]), "s": int(f.conv0.stride[0]), "p": int(f.conv0.padding[0])}
            # the torchvision stem itself (7x7 / 2 / 3 on 3 channels, 64 filters): implicit GEMM, the patch rows never reach HBM
            if tuple(w0.shape) == (64, 3, 7, 7) and self.stem_gemm["s"] == 2 and self.stem_gemm["p"] == 3:
                self.stem_gemm["w7"] = ops.stem_conv7_weights(w0, dtype)
        self.stem_pool = self._pool_spec(f.pool0, True)
        self.blocks: List[dict] = []
        i = 1
        while hasattr(f, f"denseblock{i}"):
            layers = []
            for layer in getattr(f, f"denseblock{i}").children():
                if float(getattr(layer, "drop_rate", 0.0)) > 0 and module.training:
                    raise RuntimeError("FrozenDenseNet runs the eval-mode forward")
                s1, h1 = _bn_affine(layer.norm1)
                w1, b1 = _fold(layer.conv1, layer.norm2)
                layers.append({"s1": s1, "h1": h1, "w1": cl(w1), "b1": b1.to(dtype), "w2": cl(layer.conv2.weight.detach().float()),
                               "growth": layer.conv2.out_channels, "pad2": layer.conv2.padding,
                               # GEMM form of conv1: [128, C] row-major weights, fp32 bias (None when conv1 is not 1x1 / 128 wide)
                               "w1m": (w1.reshape(w1.shape[0], -1).to(dtype).contiguous()
                                       if tuple(layer.conv1.kernel_size) == (1, 1) and w1.shape[0] == 128 else None),
                               "b1f": b1.float().contiguous()})
            blk = {"layers": layers, "transition": None}
            tr = getattr(f, f"transition{i}", None)
            if tr is not None:
                s, h = _bn_affine(tr.norm)
                blk["transition"] = {"s": s, "h": h, "w": cl(tr.conv.weight.detach().float()), "pool": self._pool_spec(tr.pool, False)}
            self.blocks.append(blk)
            i += 1
        self.s5, self.h5 = _bn_affine(f.norm5)
        self.wc = module.classifier.weight.detach().float().contiguous()
        self.bc = module.classifier.bias.detach().float().contiguous() if module.classifier.bias is not None else None
        self._fused_conv_relu = None      # decided on the first CUDA call
        self._ident = {}                  # growth -> (ones, zeros) of the identity affine used by _put_channels

    # ------------------------------------------------------------------ pieces
    @staticmethod
    def _pool_spec(p: nn.Module, is_max: bool):
        one = lambda v: v[0] if isinstance(v, (tuple, list)) else v
        if isinstance(p, nn.MaxPool2d) != is_max or not isinstance(p, (nn.MaxPool2d, nn.AvgPool2d)) or getattr(p, "ceil_mode", False):
            raise TypeError(f"unexpected pooling layer {p}")
        return (int(one(p.kernel_size)), int(one(p.stride)), int(one(p.padding)), is_max)

    @staticmethod
    def _nhwc(t):                # NCHW-logical conv output -> contiguous [B,H,W,C] (a view when the tensor is channels_last)
        t = t.permute(0, 2, 3, 1)
        return t if t.is_contiguous() else t.contiguous()

    def _conv_bias_relu(self, x, w, b, stride, padding):
        """cuDNN conv + bias + relu in one call where the build supports it for this dtype, else conv2d + in-place relu."""
        if x.is_cuda and self._fused_conv_relu is None:
            try:
                y = torch.cudnn_convolution_relu(x, w, b, list(stride), list(padding), [1, 1], 1)
                ref = F.relu(F.conv2d(x, w, b, stride, padding))
                self._fused_conv_relu = bool(y.shape == ref.shape and torch.allclose(y.float(), ref.float(), rtol=2e-2, atol=2e-2))
            except Exception:           # noqa: BLE001 -- any refusal (dtype, layout, build) selects the two-call form
                self._fused_conv_relu = False
        if x.is_cuda and self._fused_conv_relu:
            return torch.cudnn_convolution_relu(x, w, b, list(stride), list(padding), [1, 1], 1)
        return F.relu_(F.conv2d(x, w, b, stride, padding))

    def _put_channels(self, buf, c, o):
        """The layer's new features o (NCHW-logical, channels_last) -> channels [c, c + growth) of the block buffer.  ATen's
        generic strided copy needs 0.9 ms per forward for these 64-byte segments; the 16-byte-vectorised NHWC pass of
        libvitk (identity affine, no ReLU) writes them with the buffer's pixel pitch directly."""
        g = o.shape[1]
        o_nhwc = o.permute(0, 2, 3, 1)
        if buf.is_cuda and self._affine_relu is ops.affine_relu_nhwc and o_nhwc.is_contiguous() and g % 8 == 0 and c % 8 == 0:
            if g not in self._ident:
                self._ident[g] = (torch.ones(g, dtype=torch.float32, device=buf.device), torch.zeros(g, dtype=torch.float32, device=buf.device))
            one, zero = self._ident[g]
            ops.affine_relu_nhwc(o_nhwc, g, one, zero, out=buf, out_channel_offset=c, relu=False)
        else:
            buf[..., c:c + g].copy_(o_nhwc)

    @staticmethod
    def _nchw(t_nhwc):           # [B,H,W,C] contiguous -> NCHW-logical view with channels_last strides (no copy)
        return t_nhwc.permute(0, 3, 1, 2)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        x = images.to(self.dtype).contiguous(memory_format=torch.channels_last)
        w0, b0, stride0, pad0 = self.stem
        relu_after_pool = False
        sg = self.stem_gemm
        if sg is not None and self._bottleneck is not None and x.is_cuda and self.stem_pool[3]:
            xn = self._nhwc(x)                                                          # [B,H,W,3] view of the channels_last batch
            Bn, Hn, Wn, _ = xn.shape
            if "w7" in sg and self.stem_fused and Hn % 2 == 0 and Wn % 8 == 0 and Hn >= 8:
                src = ops.stem_conv7(xn, sg["w7"], sg["b"], relu=True)                 # conv0 + norm0 + relu0 (relu and max pool commute)
            else:
                patches = ops.im2col_rows(xn, sg["k"], sg["s"], sg["p"])
                OHs, OWs = (Hn + 2 * sg["p"] - sg["k"]) // sg["s"] + 1, (Wn + 2 * sg["p"] - sg["k"]) // sg["s"] + 1
                cout = sg["w"].shape[0]
                src = torch.empty(Bn, OHs, OWs, cout, dtype=self.dtype, device=x.device)
                ops.gemm(patches, sg["w"], patches.shape[0], cout, patches.shape[1], out=src, bias=sg["b"])   # conv0 + norm0, no ReLU yet
                relu_after_pool = True
        else:
            src = self._nhwc(self._conv_bias_relu(x, w0, b0, stride0, pad0))           # [B,H,W,64]
        pool = self.stem_pool
        buf = None
        for blk in self.blocks:
            B, H, W, c0 = src.shape
            k, st, pd, is_max = pool
            OH, OW = (H + 2 * pd - k) // st + 1, (W + 2 * pd - k) // st + 1
            ct = c0 + sum(l["growth"] for l in blk["layers"])
            buf = torch.empty(B, OH, OW, ct, dtype=self.dtype, device=src.device)     # the block's concatenation, NHWC
            self._pool(src, buf, k, st, pd, is_max)                                    # pooled input -> channels [0, c0)
            if relu_after_pool:                                                        # relu0 behind pool0: in place over [0, c0)
                if c0 not in self._ident:
                    self._ident[c0] = (torch.ones(c0, dtype=torch.float32, device=buf.device), torch.zeros(c0, dtype=torch.float32, device=buf.device))
                ops.affine_relu_nhwc(buf, c0, self._ident[c0][0], self._ident[c0][1], out=buf, relu=True)
                relu_after_pool = False
            c = c0
            for l in blk["layers"]:
                if self._bottleneck is not None and l["w1m"] is not None and buf.is_cuda:
                    t = self._nchw(self._bottleneck(buf, c, l["s1"], l["h1"], l["w1m"], l["b1f"]))   # one GEMM, buffer read once
                else:
                    a = self._affine_relu(buf, c, l["s1"], l["h1"])                    # norm1 + relu1 over channels [0, c)
                    t = self._conv_bias_relu(self._nchw(a), l["w1"], l["b1"], (1, 1), (0, 0))   # conv1 (+ norm2 + relu2)
                o = F.conv2d(t, l["w2"], None, 1, l["pad2"])                           # conv2: the layer's new features
                self._put_channels(buf, c, o)
                c += l["growth"]
            tr = blk["transition"]
            if tr is not None:
                a = self._affine_relu(buf, ct, tr["s"], tr["h"])
                src = self._nhwc(F.conv2d(self._nchw(a), tr["w"]))
                pool = tr["pool"]
        a = self._affine_relu(buf, buf.shape[-1], self.s5, self.h5)                    # norm5 + the forward()'s F.relu
        pooled = a.float().mean(dim=(1, 2))                                            # adaptive_avg_pool2d((1, 1)) + flatten
        return F.linear(pooled, self.wc, self.bc)

    def to(self, device):
        """Move the snapshotted parameters (used when the step is built before the module reaches the GPU)."""
        mv = lambda t: t.to(device) if isinstance(t, torch.Tensor) else t
        self.stem = tuple(mv(t) for t in self.stem)
        if self.stem_gemm is not None:
            self.stem_gemm = {k: mv(v) for k, v in self.stem_gemm.items()}
        for blk in self.blocks:
            for l in blk["layers"]:
                for k, v in l.items():
                    l[k] = mv(v)
            if blk["transition"] is not None:
                for k, v in blk["transition"].items():
                    blk["transition"][k] = mv(v)
        self.s5, self.h5, self.wc = mv(self.s5), mv(self.h5), mv(self.wc)
        self.bc = mv(self.bc)
        return self
