"""CHISCO imagined-speech epochs, post-loader contract.

The reference reads mne `.fif` epochs (data_factory/eeg_processor.py) and hands the model
`(X[B,T,C] float32, y[B], padding_mask[B,T])` batches (data_factory/eeg.py:75-93, :501-513).  mne and the
recordings are not in this environment, so ingestion starts one step later: an `.npz` / `.npy` dump of the
epochs as they leave mne — `epochs [N, C_raw, T_raw]` in volts at 500 Hz plus `labels [N]` (the 39 sentence
category ids of textmaps.json) and optional `subjects [N]` — and applies the reference's per-epoch
preprocessing and dataset logic from there on:

  * `preprocess_epoch`   = eeg_processor.py:258-376 `preprocess_eeg_data_with_downsampling`:
        decimation by q = int(original_fs / target_fs) (scipy FIR, zero phase; for the reference's fixed
        500 -> 256 Hz q is 1, scipy rejects that filter and the reference's own fallback keeps every sample),
        crop / zero-pad to 122 channels, crop / Fourier-resample to `target_timepoints`, volts -> microvolts, fp32
  * `THREE_CATEGORY_MAP` = eeg.py:63-69 `create_3category_mapping` (39 -> daily life / social-emotional /
        professional services); ids outside the map are dropped (the reference labels them -1)
  * split                = eeg.py:412-471: seeded `np.random.permutation`, 70 / 10 / 20 by default
  * sample layout        = eeg.py:501-513: `(T, C)` per item, label as a scalar
The reference's `per_sample_std` normaliser (eeg.py:332-349) groups a frame by its own row index, i.e. one row
per group, whose sample standard deviation is NaN; it is not reproduced — the shapelet expert instance-normalises
every (sample, channel) series itself (Shapelet.py:186-187), so any per-channel affine scaling cancels there.
"""
import glob
import os

import numpy as np
import torch
from torch.utils.data import Dataset

ORIGINAL_FS, TARGET_FS, TARGET_CHANNELS = 500, 256, 122        # eeg.py:142-145

_DAILY = (0, 13, 14, 18, 22, 23, 26, 35, 37)
_SOCIAL = (1, 2, 6, 7, 9, 12, 15, 17, 24, 29, 34, 36, 38)
THREE_CATEGORY_MAP = {c: (0 if c in _DAILY else 1 if c in _SOCIAL else 2) for c in range(39)}


def preprocess_epoch(epoch, target_channels=TARGET_CHANNELS, target_timepoints=None, original_fs=ORIGINAL_FS,
                     target_fs=TARGET_FS):
    """[C_raw, T_raw] volts -> [target_channels, T] microvolts fp32 (eeg_processor.py:258-376)."""
    x = np.asarray(epoch, dtype=np.float64)
    if target_fs < original_fs:
        q = int(original_fs / target_fs)
        try:
            from scipy import signal
            x = signal.decimate(x, q=q, axis=1, ftype="fir", zero_phase=True)
        except Exception:                                  # q == 1: no valid FIR design -> plain stride-q pick
            x = x[:, ::max(q, 1)]
    c, n = x.shape
    if c > target_channels:
        x = x[:target_channels]
    elif c < target_channels:
        x = np.pad(x, ((0, target_channels - c), (0, 0)))
    if target_timepoints is not None and n != target_timepoints:
        if n > target_timepoints:
            x = x[:, :target_timepoints]
        else:
            from scipy import signal
            x = signal.resample(x, target_timepoints, axis=1)
    return (x * 1e6).astype(np.float32)


def split_indices(n, flag, val_size=0.1, test_size=0.2, seed=42):
    """Seeded random train / val / test split of n samples (eeg.py:412-471)."""
    n_val, n_test = int(n * val_size), int(n * test_size)
    n_train = n - n_val - n_test
    if n_train < 1:
        n_train = 1
        n_val = min(n - 1, n_val)
        n_test = n - n_train - n_val
    elif n_val < 1 and n > 1:
        n_val = 1
        n_test = min(n - n_train - 1, n_test)
        n_train = n - n_val - n_test
    perm = np.random.RandomState(seed).permutation(n)
    return {"train": perm[:n_train], "val": perm[n_train:n_train + n_val], "test": perm[n_train + n_val:]}[flag]


class ChiscoEpochs(Dataset):
    """One split of a dumped CHISCO epoch archive.  `num_class` is 39 (`--data EEG`) or 3 (`--data EEG3`)."""

    def __init__(self, root_path, flag="train", three_class=False, max_files=None, target_timepoints=None, seed=42):
        paths = sorted(glob.glob(os.path.join(root_path, "*.npz")) + glob.glob(os.path.join(root_path, "*.npy")))
        if not paths:
            raise FileNotFoundError(f"no .npz / .npy epoch dumps under {root_path}")
        xs, ys, subj = [], [], []
        for fi, p in enumerate(paths[:max_files] if max_files else paths):
            z = np.load(p, allow_pickle=False)
            if isinstance(z, np.ndarray):
                raise ValueError(f"{p}: a bare .npy holds no labels; dump epochs, labels (and subjects) into one .npz")
            ep, lab = z["epochs"], z["labels"].astype(np.int64)
            sb = z["subjects"].astype(np.int64) if "subjects" in z.files else np.full(len(lab), fi, dtype=np.int64)
            for e, l, s in zip(ep, lab, sb):
                if three_class:
                    l = THREE_CATEGORY_MAP.get(int(l), -1)
                if l < 0:
                    continue
                xs.append(preprocess_epoch(e, target_timepoints=target_timepoints))
                ys.append(int(l)); subj.append(int(s))
        if not xs:
            raise ValueError(f"{root_path}: no usable epochs")
        keep = split_indices(len(xs), "val" if flag == "validation" else flag, seed=seed)
        self.x = torch.from_numpy(np.stack([xs[i] for i in keep], axis=0)).transpose(1, 2).contiguous()   # [N,T,C]
        self.y = torch.tensor([ys[i] for i in keep], dtype=torch.int64)
        self.subject = torch.tensor([subj[i] for i in keep], dtype=torch.int64)
        self.num_class = 3 if three_class else 39
        self.class_names = list(range(self.num_class))
        self.enc_in = int(self.x.shape[2])
        self.max_seq_len = self.seq_len = int(self.x.shape[1])

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i:i + 1]
