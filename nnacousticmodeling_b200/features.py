"""Host-side feature helpers with the reference's names (scripts/util/kw_nn_utils.py,
scripts/util/kw_utils.py).  Parsing and file formats run on the host; the per-frame arithmetic
(splice, shift, rescale, i-vector append) runs in K1 on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .engine import _device


def loadKaldiFeatureTransform(filename):
    """kw_nn_utils.py:4-11 -- nnet1 text file: <Splice> / <AddShift> / <Rescale>."""
    with open(filename) as fid:
        lines = fid.readlines()
    if len(lines) < 7:
        raise ValueError(f"{filename}: not a Kaldi nnet1 feature transform (expected >= 7 lines)")
    ft = {
        "shape": [int(t) for t in lines[1].split()[1:]],
        "shifts": [int(t) for t in lines[2].split()[1:-1]],
    }
    ft["addShift"] = np.asarray([float(t) for t in lines[4].split()[3:-1]], dtype=np.float32)
    ft["rescale"] = np.asarray([float(t) for t in lines[6].split()[3:-1]], dtype=np.float32)
    if ft["addShift"].shape != (ft["shape"][0],) or ft["rescale"].shape != (ft["shape"][0],):
        raise ValueError(f"{filename}: transform length does not match <Splice> {ft['shape']}")
    return ft


def adapt_transform(ft, network, splice, recurrent):
    """predict_folds.py:170-188 == evaluate.py:143-161: recurrent nets keep the shift-0 block, TDNN tiles
    it winlen times, feed-forward nets keep the whole vector.  Returns a new dict."""
    if ft is None:
        return None
    ft = {k: (v.copy() if isinstance(v, np.ndarray) else list(v)) for k, v in ft.items()}
    if recurrent or network == "tdnn":
        dim = ft["shape"][1]
        zi = ft["shifts"].index(0)
        mid_mul = ft["rescale"][zi * dim:(zi + 1) * dim]
        mid_add = ft["addShift"][zi * dim:(zi + 1) * dim]
        if recurrent:
            ft["rescale"], ft["addShift"] = mid_mul, mid_add
            ft["shape"][0] = dim
            ft["shifts"] = [0]
        else:
            winlen = 2 * splice + 1
            ft["rescale"], ft["addShift"] = np.tile(mid_mul, winlen), np.tile(mid_add, winlen)
            ft["shape"][0] = dim * winlen
            ft["shifts"] = list(range(-splice, splice + 1))
    return ft


def splice_and_transform(x, splice, ft=None, ivectors=None, device=0):
    """Whole-set splicing + applyKaldiFeatureTransform + i-vector concat on the device (evaluate.py:164-171).
    Returns a float32 NumPy array (N, winlen*dim [+ I]); bit-exact with the reference helpers."""
    device = _device(device)
    with torch.cuda.device(device):
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
        add = mul = iv = None
        if ft is not None:
            add = torch.from_numpy(ft["addShift"]).to(device)
            mul = torch.from_numpy(ft["rescale"]).to(device)
        if ivectors is not None:
            iv = torch.from_numpy(np.ascontiguousarray(ivectors, dtype=np.float32)).to(device)
        out, _ = ops.splice_transform(xd, xd.shape[0], splice, add, mul, iv)
        return out.cpu().numpy()


def splicing(data, iShifts, device=0):
    """kw_utils.py:24-36 for symmetric windows range(-s, s+1)."""
    sh = list(iShifts)
    s = (len(sh) - 1) // 2
    if sh != list(range(-s, s + 1)):
        raise ValueError("splicing: only symmetric contiguous windows range(-s, s+1) are supported")
    return splice_and_transform(data, s, device=device)


def saveBin(filename, x):
    """kw_utils.py:4-12 (.lab: uint32 rows, uint32 cols, row-major payload)."""
    x = np.asarray(x)
    dims = np.asarray(x.shape, dtype=np.uint32)
    if len(dims) == 1:
        dims = np.asarray([dims[0], 1], dtype=np.uint32)
    with open(filename, "wb") as fid:
        dims.tofile(fid)
        np.ascontiguousarray(x).tofile(fid)


def loadBin(filename, dtype=np.float32):
    """kw_utils.py:14-22."""
    with open(filename, "rb") as fid:
        dims = np.fromfile(fid, dtype=np.uint32, count=2)
        x = np.fromfile(fid, dtype=dtype)
    return x.reshape(dims) if dims[1] > 1 else x
