"""Fixture <-> directory tree for the reference-generated goldens under tests/golden/nets/ (TEST INFRASTRUCTURE ONLY).

A fixture is ONE ``.npz`` that holds a whole ``master_script.py``-style directory tree (generate_folds.py:98-112 /
master_script.py:63-75 layout) plus the outputs the UNMODIFIED reference produced from it
(oracle/make_golden_nets.py).  Keys:

  in:<relpath>                 array stored with np.save at <relpath>            (data_{k}.npy, offsets_{k}.npy, ...)
  npz:<relpath>::<key>         one array of the Chainer-layout archive <relpath>  (fold_{k}.npz, .../model)
  txt:<relpath>                uint8 bytes of a text file                         (test.list)
  copy:<relpath>               uint8 name of a file in tests/golden to copy       (final.feature_transform, log_ap)
  out:<name>                   golden output array (what the reference wrote)
  meta                         uint8 JSON: network description and the captured command lists

``materialise`` rebuilds the tree under a scratch root; tests then hand the captured command lists to the product's
``predict.main`` / ``evaluate.main`` from inside that root (all paths in the lists are relative to it).
"""
import json
import os
import shutil

import numpy as np


def pack_text(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def unpack_text(a):
    return bytes(np.asarray(a, dtype=np.uint8)).decode()


def materialise(fixture, root, golden_dir):
    """Write the input tree of ``fixture`` (path of the .npz) under ``root``; returns the ``meta`` dict."""
    z = np.load(fixture)
    archives = {}
    for key in z.files:
        kind, _, rest = key.partition(":")
        if kind == "npz":
            rel, _, k = rest.partition("::")
            archives.setdefault(rel, {})[k] = z[key].astype(np.float32)  # stored as (exact) fp16
            continue
        if kind not in ("in", "txt", "copy"):
            continue
        path = os.path.join(root, rest)
        os.makedirs(os.path.dirname(path) or root, exist_ok=True)
        if kind == "in":
            np.save(path, z[key])
        elif kind == "txt":
            with open(path, "w") as f:
                f.write(unpack_text(z[key]))
        else:
            shutil.copy(os.path.join(golden_dir, unpack_text(z[key])), path)
    for rel, arrays in archives.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:  # a bare file name such as ".../model" must not grow a ".npz" suffix
            np.savez(f, **arrays)
    return json.loads(unpack_text(z["meta"]))


def golden_outputs(fixture):
    z = np.load(fixture)
    return {k.partition(":")[2]: z[k] for k in z.files if k.startswith("out:")}


def archive_params(fixture, rel):
    """The parameters of archive ``rel`` as {key without 'predictor/': array} (the oracle's parameter dicts)."""
    z = np.load(fixture)
    pre = f"npz:{rel}::"
    return {k[len(pre):].replace("predictor/", "", 1): z[k].astype(np.float32) for k in z.files if k.startswith(pre)}


def load_lab(path):
    """kw_utils.py:14-22 (loadBin): uint32 rows, uint32 cols, float32 row-major."""
    with open(path, "rb") as f:
        dims = np.fromfile(f, dtype=np.uint32, count=2)
        return np.fromfile(f, dtype=np.float32).reshape(dims)
