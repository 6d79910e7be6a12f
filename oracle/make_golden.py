#!/usr/bin/env python3
"""Generate tests/golden/* by running the REFERENCE's own NumPy helpers.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python oracle/make_golden.py

It imports, unmodified, from /root/reference/scripts/util:
  kw_utils.splicing / logsum / saveBin / loadBin          (kw_utils.py:4-43)
  kw_nn_utils.loadKaldiFeatureTransform / applyKaldiFeatureTransform / prepareBatch
                                                          (kw_nn_utils.py:4-43)
  orcus_util.apply_time_delay                             (orcus_util.py:13-43)
and stores their inputs and outputs as small .npz fixtures.  It also copies the two DATA
fixtures the path reads (kaldi/final.feature_transform, recog/log_ap_Kaldi1909.npy) plus
kaldi/ali_train_pdf.counts so that the parser and the prior can be checked without the
reference checkout.  No reference source code is copied.
"""
import os
import shutil
import sys

import numpy as np

REF = os.environ.get("NNAM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(REF, "scripts", "util"))

from kw_utils import splicing, logsum, saveBin, loadBin  # noqa: E402
from kw_nn_utils import loadKaldiFeatureTransform, applyKaldiFeatureTransform, prepareBatch  # noqa: E402
from orcus_util import apply_time_delay  # noqa: E402


def main():
    os.makedirs(OUT, exist_ok=True)
    for src in ("kaldi/final.feature_transform", "recog/log_ap_Kaldi1909.npy", "kaldi/ali_train_pdf.counts"):
        dst = os.path.join(OUT, os.path.basename(src))
        shutil.copyfile(os.path.join(REF, src), dst)
        os.chmod(dst, 0o644)

    rng = np.random.default_rng(20261018)
    ft = loadKaldiFeatureTransform(os.path.join(REF, "kaldi", "final.feature_transform"))
    np.savez(os.path.join(OUT, "feature_transform_parsed.npz"),
             shape=np.asarray(ft["shape"]), shifts=np.asarray(ft["shifts"]),
             addShift=ft["addShift"], rescale=ft["rescale"])

    # --- splice: two independent reference implementations + transform ------------------
    x = rng.standard_normal((257, 40)).astype(np.float32)
    sp = splicing(x, range(-5, 6))
    pb, _ = prepareBatch(x, [], np.arange(x.shape[0]), 11)
    assert np.array_equal(sp, pb)
    idx = np.array([0, 1, 4, 5, 6, 100, 250, 251, 252, 255, 256])
    pb_sub, _ = prepareBatch(x, [], idx.copy(), 11)
    sp_ft = applyKaldiFeatureTransform(sp, ft)
    # tiny array: window wider than the data (both clamps active in one row)
    xs = rng.standard_normal((3, 40)).astype(np.float32)
    sp_small = splicing(xs, range(-5, 6))
    # other splice widths used by the example scripts (TDNN: ksize 5,5,5,5 -> splice 8)
    sp17 = splicing(x[:64], range(-8, 9))
    sp1 = splicing(x[:8], range(0, 1))
    np.savez(os.path.join(OUT, "splice.npz"), x=x, splice11=sp, idx=idx, prepare_sub=pb_sub,
             splice11_ft=sp_ft, x_small=xs, splice11_small=sp_small, splice17=sp17, splice1=sp1)

    # --- recurrent transform = middle block, time delay -------------------------------------
    dim = ft["shape"][1]
    zi = ft["shifts"].index(0)
    ftm = {"addShift": ft["addShift"][zi * dim:(zi + 1) * dim], "rescale": ft["rescale"][zi * dim:(zi + 1) * dim]}
    offsets = np.array([0, 7, 19, 20, 64], dtype=np.int32)
    xr = rng.standard_normal((64, 40)).astype(np.float32)
    xd, _, offd = apply_time_delay(xr, None, offsets, 5)
    np.savez(os.path.join(OUT, "timedelay.npz"), x=xr, offsets=offsets, x_delayed=xd, offsets_delayed=offd,
             x_mid_ft=applyKaldiFeatureTransform(xr, ftm))

    # --- logsum / head ------------------------------------------------------------------
    ap = np.load(os.path.join(REF, "recog", "log_ap_Kaldi1909.npy"))
    y = (3.0 * rng.standard_normal((33, 1909))).astype(np.float32)
    y[5] = -1e4
    y[6, 3] = 80.0
    ls = logsum(y, axis=1)
    ya = y - ap
    np.savez(os.path.join(OUT, "head.npz"), y=y, logsum=ls, logsoftmax=y - ls,
             head_ap=ya - logsum(ya, axis=1))

    # --- .lab format --------------------------------------------------------------------
    lab = os.path.join(OUT, "sample.lab")
    saveBin(lab, (y - ls)[:4])
    back = loadBin(lab, np.float32)
    assert np.array_equal(back, (y - ls)[:4])
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
