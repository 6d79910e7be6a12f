"""TDNN (chainer_networks.py:24-42) as K2 GEMMs -- "CNN-as-GEMM" with no im2col copy.

Reference: the flat (B, win*D) input row is reshaped to (B, C, 1, win) WITHOUT a transpose (quirk Q3:
flat index n is read as channel c = n // win, position w = n % win), then L valid 1 x k convolutions
(+ activation), then ``out`` Linear on the remaining (B, units_last) -- the window shrinks to width 1.

Device layout: the activations of layer l are one row per frame, ``(rows, W_l * C_l)`` with the window
position major and the channel minor.  A valid 1 x k convolution evaluated at position w reads the
CONTIGUOUS span ``[w*C, (w+k)*C)`` of that row, so it is the plain Linear kernel with a pointer offset:

    out[:, w*U:(w+1)*U] = act( in[:, w*C:(w+k)*C] . W2^T + b ),   W2[o, j*C + c] = W[o, c, 0, j]

one K2 launch per output position, no data movement.  The first layer has to undo the scrambled input
layout instead: it is ONE GEMM on the flat spliced row with the kernel scattered into a
(W_1*U_0, C_0*win) matrix (zeros elsewhere): Wfull[w*U + o, c*win + w + j] = W[o, c, 0, j].  That spends
win/k times the first layer's (small) FLOPs and keeps K1's output usable as is.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._native import NnamError
from .ops import OUT_F32, round_up


def build_plan(plan, model):
    from .engine import LinearDev
    p, dev, split = model.params, plan.device, plan.split
    win = model.input_win_size
    w0 = p["layer_0/W"]
    u0, c0, _, k0 = w0.shape
    if any(u % 8 for u in model.n_units):
        raise NnamError("tdnn: unit counts must be multiples of 8 on the B200 path (16-byte aligned row slices)")
    widths = [win]
    for k in model.ksize:
        widths.append(widths[-1] - k + 1)
    if widths[-1] != 1:
        raise NnamError("tdnn: the kernel sizes must reduce the input window to width 1")
    plan.tdnn_widths = widths
    w1 = widths[1]
    full = np.zeros((w1 * u0, c0 * win), dtype=np.float32)
    for w in range(w1):
        for j in range(k0):
            full[w * u0:(w + 1) * u0, w + j::win] = w0[:, :, 0, j]
    plan.tdnn_first = LinearDev(full, np.tile(p["layer_0/b"], w1), dev, split, plan.elem)
    plan.tdnn_convs = []
    for l in range(1, model.layers):
        w = p[f"layer_{l}/W"]  # (out, in, 1, k)
        w2 = np.ascontiguousarray(np.transpose(w[:, :, 0, :], (0, 2, 1))).reshape(w.shape[0], -1)  # (out, k*in)
        plan.tdnn_convs.append(LinearDev(w2, p[f"layer_{l}/b"], dev, split, plan.elem))
    plan.out = LinearDev(p["out/W"], p["out/b"], dev, split, plan.elem)


def logits(model, plan, a_hi, a_lo, rows, tag="tdnn", ws=None):
    """Conv stack + ``out`` on staged bf16 inputs (rows, roundup(C*win, 8)); returns fp32 logits in workspace."""
    ws, act, kind = ws or plan.ws, model.activation.name, plan.act_kind
    cap = a_hi.shape[0]
    widths, units = plan.tdnn_widths, model.n_units

    def buf(l):
        ld = widths[l + 1] * units[l]
        hi = ws.get(f"act.h{l % 2}.hi", cap, ld, plan.tdt)
        lo = ws.get(f"act.h{l % 2}.lo", cap, ld, torch.bfloat16) if plan.split else None
        return hi, lo

    hi, lo = buf(0)
    plan.tdnn_first(a_hi, a_lo, rows, act, kind, out=(hi, lo))
    for l in range(1, model.layers):
        lin = plan.tdnn_convs[l - 1]
        c, u = units[l - 1], units[l]
        nhi, nlo = buf(l)
        for w in range(widths[l + 1]):
            ops.linear_bias_act(hi[:, w * c:], None if lo is None else lo[:, w * c:], lin.w_hi, lin.w_lo, lin.bias,
                                rows, lin.n, lin.k, act=act, out_kind=kind, nsplit=3 if plan.split else 1,
                                out=(nhi[:, w * u:], None if nlo is None else nlo[:, w * u:]), ldo=nhi.stride(0),
                                elem=plan.elem)
        hi, lo = nhi, nlo
    out = ws.get(f"{tag}.logits", cap, round_up(plan.out.n, 16), torch.float32)
    plan.out(hi, lo, rows, "identity", OUT_F32, out=(out, None))
    return out
