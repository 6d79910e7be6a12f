"""Synthetic TIMIT-shaped data sets in the reference's file convention (generate_folds.py:98-112:
``data_{k}.npy`` (N, 40) float32, ``offsets_{k}.npy`` (U+1,) int32 with a leading 0, ``ivectors_{k}.npy`` (N, I)
float32).  There is no network for the real corpus, so benchmarks and examples run on these; the generator is
deterministic per seed."""
from __future__ import annotations

import numpy as np

TIMIT_TRAIN_UTTS, TIMIT_TRAIN_FRAMES = 3696, 1124823  # kaldi/ali_train_pdf.counts; scripts/common/predict_folds.py:139
TIMIT_TEST_SHAPED_UTTS = 1344                         # BASELINE.json configs[0]


def synth_lengths(rng, n_utt, total=None):
    """Utterance lengths: lognormal(ln 295, 0.28) clipped to [90, 780] frames (mean ~305); optionally adjusted, 16
    frames at a time over the utterances in order, until they sum to ``total``."""
    ln = np.clip(np.round(rng.lognormal(np.log(295.0), 0.28, n_utt)), 90, 780).astype(np.int64)
    if total is not None:
        diff = int(total - ln.sum())
        i = 0
        while diff != 0:
            step = int(np.sign(diff)) * min(abs(diff), 16)
            new = int(np.clip(ln[i % n_utt] + step, 90, 780))
            diff -= new - ln[i % n_utt]
            ln[i % n_utt] = new
            i += 1
    return ln


def synth_set(seed, n_utt, dim=40, ivec_dim=0, total=None, utts_per_speaker=8):
    """(x, offsets, ivectors): unit-variance features, one N(0, 0.5^2) i-vector per synthetic speaker repeated on every
    frame of the speaker's utterances (offline per-speaker i-vectors), or None."""
    rng = np.random.default_rng(seed)
    ln = synth_lengths(rng, n_utt, total)
    offsets = np.concatenate([[0], np.cumsum(ln)]).astype(np.int32)
    n = int(offsets[-1])
    x = rng.standard_normal((n, dim), dtype=np.float32)
    iv = None
    if ivec_dim:
        n_spk = (n_utt + utts_per_speaker - 1) // utts_per_speaker
        spk = (0.5 * rng.standard_normal((n_spk, ivec_dim))).astype(np.float32)
        iv = np.repeat(spk[np.arange(n_utt) // utts_per_speaker], ln, axis=0)
    return x, offsets, iv
