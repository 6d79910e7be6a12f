"""PeepholeLSTM (chainer_networks.py:103-121, L.StatefulPeepholeLSTM) on the device, time step by time step.

The peephole cell needs three extra H x H products per step (P_i c, P_f c before the cell update, P_o c' after it),
so one unit slice is 7 H-long weight rows and does not fit in shared memory next to the operand tiles of the
persistent K3 kernel at H = 512.  Until it gets its own persistent variant it runs on the packed time-major rows with
four launches per step and layer:

    g1 = [h | c]_{t-1} . [lateral/W | P]^T        K2 (N = 4H, K = 2H; P rows [0, peep_i, peep_f, 0] per unit)
    c' = cell phase 0 (gx_t, g1, c_{t-1})          peephole_cell_kernel<0>
    p2 = c' . peep_o^T                             K2 (N = H, K = H; A = the c columns of this step's [h | c] rows)
    h' = cell phase 1 (gx_t, g1, p2, c')           peephole_cell_kernel<1>

The input projection of all frames is one GEMM per layer, as for LSTM/GRU.  Rows are stored as [h | c] bf16 (hi, lo)
pairs so that both products read their A operand in place; the next layer reads the h half through the leading
dimension.  Launch-bound (about 16 launches per time step for 4 layers) -- it exists for coverage and parity.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .engine import LinearDev
from .ops import OUT_F32


class PeepLayer:
    pass


def build_plan(plan, model):
    p, dev, split, h = model.params, plan.device, plan.split, model.n_units
    plan.hidden = h
    plan.n_dirs = 1
    plan.peep_layers = []
    for l in range(model.layers):
        pre = f"layer_{l}/"
        L = PeepLayer()
        L.upward = LinearDev(p[pre + "upward/W"], p[pre + "upward/b"], dev, split, plan.elem)
        pmat = np.zeros((4 * h, h), np.float32)  # gate-interleaved rows [a, i, f, o] per unit
        pmat[1::4] = p[pre + "peep_i/W"]
        pmat[2::4] = p[pre + "peep_f/W"]
        L.w1 = LinearDev(np.concatenate([p[pre + "lateral/W"], pmat], axis=1), None, dev, split, plan.elem)
        L.w2 = LinearDev(p[pre + "peep_o/W"], None, dev, split, plan.elem)
        plan.peep_layers.append(L)
    plan.out = LinearDev(p["out/W"], p["out/b"], dev, split, plan.elem)


def _step(plan, layer, ws, gx_t, hc_prev, c_prev, hc_t, c_t, n):
    """One time step of one layer on n packed rows.  hc_*: (hi, lo) views of [h | c] rows; c_*: fp32 (n, H)."""
    h = plan.hidden
    nsplit = 3 if plan.split else 1
    g1 = None
    if hc_prev is not None:
        g1 = ws.get("peep.g1", n, 4 * h, torch.float32)
        ops.linear_bias_act(hc_prev[0], hc_prev[1], layer.w1.w_hi, layer.w1.w_lo, None, n, 4 * h, 2 * h,
                            out_kind=OUT_F32, nsplit=nsplit, out=(g1, None), elem=plan.elem)
    ops.peephole_cell(0, gx_t, g1, None, c_prev, c_t, hc_t[0], hc_t[1], n, h, not plan.split, plan.elem)
    p2 = ws.get("peep.p2", n, h, torch.float32)
    ops.linear_bias_act(hc_t[0][:, h:], None if hc_t[1] is None else hc_t[1][:, h:], layer.w2.w_hi, layer.w2.w_lo, None,
                        n, h, h, out_kind=OUT_F32, nsplit=nsplit, out=(p2, None), elem=plan.elem)
    ops.peephole_cell(1, gx_t, g1, p2, None, c_t, hc_t[0], hc_t[1], n, h, not plan.split, plan.elem)


def run_layers(model, plan, sched, a_hi, a_lo, rows, ws=None):
    """All peephole layers on the packed rows of a ONE-batch schedule; returns (h_hi, h_lo) views whose first H
    columns are the last layer's output (leading dimension 2H)."""
    ws = ws or plan.ws
    h = plan.hidden
    base = sched.h_base[0]  # prefix sums of the active counts, steps + 1 entries
    for l, layer in enumerate(plan.peep_layers):
        gx = ws.get("peep.gx", rows, 4 * h, torch.float32)
        layer.upward(a_hi, a_lo, rows, "identity", OUT_F32, out=(gx, None))
        hc_hi = ws.get(f"peep.hc{l % 2}.hi", rows, 2 * h, plan.tdt)
        hc_lo = ws.get(f"peep.hc{l % 2}.lo", rows, 2 * h, torch.bfloat16) if plan.split else None
        c = ws.get("peep.c", rows, h, torch.float32)
        for t in range(len(base) - 1):
            r0, r1 = int(base[t]), int(base[t + 1])
            n = r1 - r0
            prev = None
            c_prev = None
            if t > 0:
                q0 = int(base[t - 1])
                prev = (hc_hi[q0:q0 + n], None if hc_lo is None else hc_lo[q0:q0 + n])
                c_prev = c[q0:q0 + n]
            _step(plan, layer, ws, gx[r0:r1], prev, c_prev, (hc_hi[r0:r1], None if hc_lo is None else hc_lo[r0:r1]),
                  c[r0:r1], n)
        a_hi, a_lo = hc_hi, hc_lo
    return a_hi, a_lo


def step(model, plan, xd):
    """Stateful ``model(x)``: one time step for a batch of B rows (chainer_networks.py:113-121)."""
    from .engine import round_up
    from ._native import NnamError
    B, h = xd.shape[0], plan.hidden
    st = model._state
    if st is not None and st["B"] != B:
        raise NnamError("model(x): batch size changed between steps; call reset_state() first")
    a_hi, a_lo = ops.convert_f32(xd.contiguous(), plan.act_kind)
    new = []
    for l, layer in enumerate(plan.peep_layers):
        gx = plan.ws.get("peepstep.gx", B, 4 * h, torch.float32)
        layer.upward(a_hi, a_lo, B, "identity", OUT_F32, out=(gx, None))
        hc_hi = torch.empty((B, 2 * h), dtype=plan.tdt, device=plan.device)
        hc_lo = torch.empty((B, 2 * h), dtype=torch.bfloat16, device=plan.device) if plan.split else None
        c = torch.empty((B, h), dtype=torch.float32, device=plan.device)
        prev, c_prev = (None, None) if st is None else st["s"][l]
        _step(plan, layer, plan.ws, gx, prev, c_prev, (hc_hi, hc_lo), c, B)
        new.append(((hc_hi, hc_lo), c))
        a_hi, a_lo = hc_hi, hc_lo
    model._state = {"B": B, "s": new}
    logits = plan.ws.get("peepstep.logits", B, round_up(model.n_out, 16), torch.float32)
    plan.out(a_hi, a_lo, B, "identity", OUT_F32, out=(logits, None))
    return logits[:B, :model.n_out].clone()
