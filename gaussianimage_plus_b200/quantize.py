"""Attribute quantisers of the compression pass (`train_quantize.py`): host-side mirror of
the reference's default `lsq` path -- `UniformQuantizer` (quantize.py:39-156), `LogQuantizer` (:158-258) and
`HybirdQuant` (:336-389) -- with the same class names, constructor arguments, `forward` return tuple
`(dequant, entropy_loss, bits, code)`, `compress` / `decompress` / `size` / `reset_state`, and the same
`scale` / `beta` parameter names (state-dict compatible).

What is different underneath: the reference builds each quantiser out of ~10 elementwise autograd nodes
(`grad_scale`, `clamp`, `ste`, ...); here each one is ONE autograd function with the backward written out --

  LSQ       code = (x - beta) / s,  c = clamp(code, qmin, qm),  r = round(c),  y = r s + beta
            dy/dx = [qmin <= code <= qm],  dy/ds = r - [..] code,  dy/dbeta = 1 - [..]        (summed over N)
  log       L = log(|x| + 1e-6),  beta = min L,  M = max L  (over the whole tensor, with their gradients, as
            quantize.py:225-229 recomputes them every call),  s = (M - beta) / (qmax - qmin),
            y = exp(round(clamp((L - beta) / s)) s + beta)

-- so a quantised forward costs a handful of launches instead of ~60 and the three quantisers sit directly in
front of the projection kernel.  These are O(N) elementwise ops on 8 floats per Gaussian; the rasterization
they feed is where the time goes.  The VQ (`vector_quantize_pytorch`) and entropy-coding (`constriction`)
options of the reference are out of scope (third-party packages that are absent here).
"""
from __future__ import annotations

import torch
from torch import nn


class _LsqFn(torch.autograd.Function):
    """y = round(clamp((x - beta) / s, qmin, qm)) * s + beta with the straight-through estimator."""

    @staticmethod
    def forward(ctx, x, scale, beta, qmin, qm):
        code = (x - beta) / scale
        inside = (code >= qmin) & (code <= qm)          # torch.clamp passes the gradient on the closed interval
        r = code.clamp(qmin, qm).round()
        ctx.save_for_backward(code, inside, r)
        ctx.mark_non_differentiable(r)
        return r * scale + beta, r

    @staticmethod
    def backward(ctx, g, _g_code):
        code, inside, r = ctx.saved_tensors
        m = inside.to(g.dtype)
        gx = g * m
        gs = (g * (r - m * code)).sum(dim=0)
        gb = (g * (1 - m)).sum(dim=0)
        return gx, gs, gb, None, None


class UniformQuantizer(nn.Module):
    """LSQ+ (quantize.py:39-156): per-channel learned step `scale` and offset `beta`, initialised from the
    first batch's min / max (`_init_data`, :72-80)."""

    def __init__(self, signed=False, bits=8, learned=False, num_channels=1, entropy_type="none", weight=0.0001):
        super().__init__()
        self.bits = bits
        self.init_state = 0
        if signed:
            self.qmin, self.qmax = -2 ** (bits - 1), 2 ** (bits - 1) - 1
        else:
            self.qmin, self.qmax = 0, 2 ** bits - 1
        self.qm = self.qmax
        self.learned = learned
        self.entropy_type = entropy_type
        if learned:
            self.scale = nn.Parameter(torch.ones(num_channels) / self.qmax)
            self.beta = nn.Parameter(torch.ones(num_channels) / self.qmax)

    def _init_data(self, tensor):
        t_min, t_max = tensor.min(dim=0)[0], tensor.max(dim=0)[0]
        scale = (t_max - t_min) / (self.qmax - self.qmin)
        self.beta.data = (t_min - self.qmin * scale).detach().to(tensor.device)
        self.scale.data = scale.detach().to(tensor.device)

    def forward(self, x, quant_loss=False):
        if self.init_state == 0:
            self._init_data(x)
            self.init_state += 1
        dequant, code = _LsqFn.apply(x, self.scale, self.beta, self.qmin, self.qm)
        return dequant, 0, 0, code

    def size(self):
        return self.bits

    def reset_state(self):
        self.init_state = 0

    def compress(self, x):
        code = ((x - self.beta) / self.scale).clamp(self.qmin, self.qmax).round()
        return code * self.scale + self.beta, code

    def decompress(self, x):
        return x * self.scale + self.beta


class _LogQuantFn(torch.autograd.Function):
    """The non-learned branch of LogQuantizer.forward (quantize.py:223-235), backward through min / max too."""

    @staticmethod
    def forward(ctx, x, qmin, qmax):
        L = torch.log(torch.abs(x) + 1e-6)
        beta, M = L.min(), L.max()
        scale = (M - beta) / (qmax - qmin)
        code = (L - beta) / scale
        inside = (code >= qmin) & (code <= qmax)
        r = code.clamp(qmin, qmax).round()
        y = torch.exp(r * scale + beta)
        ctx.save_for_backward(x, L, code, inside, r, y, beta, M)
        ctx.span = float(qmax - qmin)
        ctx.mark_non_differentiable(r, beta, scale, M)
        return y, r, beta, scale, M

    @staticmethod
    def backward(ctx, g, *_):
        x, L, code, inside, r, y, beta, M = ctx.saved_tensors
        m = inside.to(g.dtype)
        gq = g * y                                      # through exp
        g_scale = (gq * (r - m * code)).sum()
        g_beta = (gq * (1 - m)).sum() - g_scale / ctx.span
        g_max = g_scale / ctx.span
        at_min, at_max = (L == beta), (L == M)          # torch.min()/max() spread the gradient over ties
        gL = gq * m + at_min * (g_beta / at_min.sum()) + at_max * (g_max / at_max.sum())
        return gL * torch.sign(x) / (torch.abs(x) + 1e-6), None, None


class LogQuantizer(nn.Module):
    """quantize.py:158-258.  Only the configuration the live model uses is mirrored: `learned=False`
    (HybirdQuant builds it that way, :344), where range and step come from the data on every call."""

    def __init__(self, signed=True, bits=8, learned=False, num_channels=1, entropy_type="none", weight=0.001):
        super().__init__()
        if learned:
            raise NotImplementedError("the reference never instantiates a learned LogQuantizer")
        self.bits = bits
        self.init_state = 0
        if signed:
            self.qmin, self.qmax = -2 ** (bits - 1), 2 ** (bits - 1) - 1
        else:
            self.qmin, self.qmax = 0, 2 ** bits - 1
        self.learned = False
        self.beta = torch.empty(num_channels)
        self.scale = torch.empty(num_channels)
        self.min_log, self.max_log = 0, 0
        self.sign = None

    def _init_data(self, tensor):
        L = torch.log(torch.abs(tensor) + 1e-6)
        t_min, t_max = L.min(dim=0)[0], L.max(dim=0)[0]
        self.scale = ((t_max - t_min) / (self.qmax - self.qmin)).detach()
        self.beta = t_min.detach()
        self.min_log, self.max_log = self.beta, t_max.detach()

    def forward(self, x, quant_loss=False):
        if self.init_state == 0:
            self._init_data(x)
            self.init_state += 1
        y, code, beta, scale, M = _LogQuantFn.apply(x, self.qmin, self.qmax)
        self.beta, self.scale, self.max_log = beta, scale, M   # (the reference keeps the per-call values, :226-229)
        return y, 0, 0, code

    def size(self):
        return self.bits

    def reset_state(self):
        self.init_state = 0

    def compress(self, x):
        self._init_data(x)                                      # per-channel range (:244-245)
        L = torch.log(torch.abs(x) + 1e-6)
        code = ((L - self.beta) / self.scale).clamp(self.qmin, self.qmax).round()
        self.sign = torch.sign(x)
        return torch.exp(code * self.scale + self.beta), code

    def decompress(self, x):
        return torch.exp(x * self.scale + self.beta)


class HybirdQuant(nn.Module):
    """quantize.py:336-389: the two variances (columns 0, 2) through the log quantiser, the covariance
    (column 1) through a learned LSQ quantiser."""

    def __init__(self, signed=False, bits=8, cov_bits=10, learned=False, num_channels=1, entropy_type="none",
                 weight=0.001):
        super().__init__()
        self.init_state = 0
        self.var_quantizer = LogQuantizer(False, bits, learned=False, num_channels=2, entropy_type=entropy_type,
                                          weight=weight)
        self.cov_quantizer = UniformQuantizer(signed, cov_bits, learned=True, num_channels=1,
                                              entropy_type=entropy_type, weight=weight)
        self.bits = bits

    def _init_data(self, tensor):
        self.var_quantizer._init_data(tensor[:, ::2])
        self.cov_quantizer._init_data(tensor[:, 1:2])

    def forward(self, x, quant_loss=False):
        if self.init_state == 0:
            self._init_data(x)
            self.init_state += 1
        dv, _, _, cv = self.var_quantizer(x[:, ::2], quant_loss)
        dc, _, _, cc = self.cov_quantizer(x[:, 1:2], quant_loss)
        return (torch.cat([dv[:, 0:1], dc, dv[:, 1:2]], dim=1), 0, 0,
                torch.cat([cv[:, 0:1], cc, cv[:, 1:]], dim=1))

    def size(self):
        return (self.cov_quantizer.size() + self.var_quantizer.size() * 2) / 3

    def reset_state(self):
        self.var_quantizer.reset_state()
        self.cov_quantizer.reset_state()

    def compress(self, x):
        dv, cv = self.var_quantizer.compress(x[:, ::2])
        dc, cc = self.cov_quantizer.compress(x[:, 1:2])
        return (torch.cat([dv[:, 0:1], dc, dv[:, 1:2]], dim=1), torch.cat([cv[:, 0:1], cc, cv[:, 1:]], dim=1))

    def decompress(self, x):
        var = self.var_quantizer.decompress(x[:, ::2])
        cov = self.cov_quantizer.decompress(x[:, 1:2])
        return torch.cat([var[:, 0:1], cov, var[:, 1:2]], dim=1)
