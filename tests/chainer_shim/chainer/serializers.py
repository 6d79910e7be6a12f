"""chainer.serializers.{load,save}_npz: keys are the ``namedparams`` paths without the leading slash
(chainer/serializers/npz.py: DictionarySerializer joins the link hierarchy with '/'), e.g.
``predictor/layer_0/upward/W``.  Loading fills uninitialised parameters with the stored shape (lazy ``in_size``)
and is strict: a parameter missing from the file is a KeyError, a shape mismatch a ValueError."""
import numpy as np


def save_npz(file, obj, compression=True):
    arrays = {path.lstrip("/"): p.data for path, p in obj.namedparams() if p.data is not None}
    with open(file, "wb") as f:  # np.savez would append '.npz' to a bare path such as ".../model"
        (np.savez_compressed if compression else np.savez)(f, **arrays)


def load_npz(file, obj, path="", strict=True):
    with np.load(str(file)) as f:
        for name, p in obj.namedparams():
            key = path + name.lstrip("/")
            if key not in f.files:
                if strict:
                    raise KeyError(f"{key} is not in the npz file {file}")
                continue
            value = np.asarray(f[key])
            if p.data is not None and p.data.shape != value.shape:
                raise ValueError(f"{key}: shape mismatch {p.data.shape} vs {value.shape}")
            p.data = value.astype(np.float32, copy=True)
