"""UEA / UCR `.ts` archives without sktime: a from-scratch reader of the published `.ts` text format
plus the reference loader's preprocessing (reference data_factory/data_loader.py:600-719, uea.py:7-125):

  * `<root>/<name>_TRAIN.ts` for flag 'train', `<name>_TEST.ts` for 'val' / 'test' (the TSLib convention the reference
    inherits; its own case-sensitive `re.search(flag, path)` never matches the upper-case file names)
  * class labels -> sorted categories -> integer codes, returned as y[1]
  * '?' / NaN samples: linear interpolation in both directions per series and dimension (`interpolate_missing`)
  * series longer than 256 in archives whose dimensions differ in length are subsampled by 2 (`subsample`)
  * standardisation across ALL rows of the split per feature: (x - mean) / (std(ddof=1) + eps)   (`Normalizer`)
  * batches: zero padding / clipping to `max_len` and a boolean padding mask (1 = keep)           (`collate_fn`)

The `.ts` format: header lines `@key value...` (`@problemName`, `@timeStamps`, `@missing`, `@univariate`,
`@dimensions`, `@equalLength`, `@seriesLength`, `@classLabel true a b c`), then `@data`, then one case per line:
dimensions separated by ':', samples by ',', the class value last.
"""
import glob
import os

import numpy as np
import torch
from torch.utils.data import Dataset


def read_ts(path):
    """-> (cases, labels, meta): cases[i] is a list of 1-D float64 arrays (one per dimension, NaN for '?')."""
    meta, cases, labels = {}, [], []
    in_data = False
    with open(path, "r", encoding="utf-8") as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            if not in_data:
                if line.lower().startswith("@data"):
                    in_data = True
                elif line.startswith("@"):
                    parts = line[1:].split()
                    meta[parts[0].lower()] = parts[1:]
                continue
            fields = line.split(":")
            has_label = meta.get("classlabel", ["false"])[0].lower() == "true"
            dims = fields[:-1] if has_label else fields
            if has_label:
                labels.append(fields[-1].strip())
            series = []
            for d in dims:
                vals = [v.strip() for v in d.split(",") if v.strip() != ""]
                series.append(np.array([np.nan if v in ("?", "NaN", "nan") else float(v) for v in vals], dtype=np.float64))
            cases.append(series)
    if not cases:
        raise ValueError(f"{path}: no cases after @data")
    return cases, labels, meta


def _interpolate(y):
    """Linear interpolation of NaNs, extended with the edge values (pandas interpolate(limit_direction='both'))."""
    bad = np.isnan(y)
    if not bad.any():
        return y
    good = np.nonzero(~bad)[0]
    if good.size == 0:
        return y
    out = y.copy()
    out[bad] = np.interp(np.nonzero(bad)[0], good, y[good])
    return out


class UEADataset(Dataset):
    """One split of a UEA archive with the reference loader's preprocessing; `.max_seq_len`, `.enc_in`,
    `.num_class`, `.class_names` are what Experiment._get_params_from_data reads."""

    def __init__(self, root_path, flag="train", limit_size=None):
        want = "_TRAIN.ts" if flag.lower() == "train" else "_TEST.ts"
        paths = sorted(p for p in glob.glob(os.path.join(root_path, "*")) if p.upper().endswith(want.upper()))
        if not paths:
            raise FileNotFoundError(f"no *{want} file under {root_path}")
        cases, labels, self.meta = read_ts(paths[0])
        self.class_names = sorted(set(labels))
        code = {c: i for i, c in enumerate(self.class_names)}
        self.y = torch.tensor([code[l] for l in labels], dtype=torch.int64)
        lens = np.array([[len(s) for s in case] for case in cases])
        if np.abs(lens - lens[:, :1]).sum() > 0:                     # dimensions of different length: subsample long ones
            cases = [[s[::2] if len(s) > 256 else s for s in case] for case in cases]
            lens = np.array([[len(s) for s in case] for case in cases])
        self.max_seq_len = int(lens[:, 0].max())
        xs = []
        for case, ln in zip(cases, lens[:, 0]):
            xs.append(np.stack([_interpolate(s[:ln] if len(s) >= ln else np.pad(s, (0, ln - len(s)), constant_values=np.nan))
                                for s in case], axis=1))             # [T_i, C]
        if limit_size is not None:
            n = int(limit_size) if limit_size > 1 else int(limit_size * len(xs))
            xs, self.y = xs[:n], self.y[:n]
        allrows = np.concatenate(xs, axis=0)
        self.mean = np.nanmean(allrows, axis=0)
        self.std = np.nanstd(allrows, axis=0, ddof=1)
        eps = np.finfo(float).eps
        self.x = [torch.from_numpy(((x - self.mean) / (self.std + eps)).astype(np.float32)) for x in xs]
        self.enc_in = int(allrows.shape[1])
        self.num_class = len(self.class_names)
        self.seq_len = self.max_seq_len

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i], self.y[i:i + 1]


def collate_fn(batch, max_len=None):
    """(x[T_i,C], y[1]) samples -> (X[B,max_len,C] zero padded / clipped, y[B,1], padding_mask[B,max_len] bool)."""
    xs, ys = zip(*batch)
    lengths = [x.shape[0] for x in xs]
    if max_len is None:
        max_len = max(lengths)
    X = torch.zeros(len(xs), max_len, xs[0].shape[-1])
    for i, x in enumerate(xs):
        end = min(lengths[i], max_len)
        X[i, :end] = x[:end]
    mask = torch.arange(max_len).unsqueeze(0) < torch.tensor(lengths).clamp(max=max_len).unsqueeze(1)
    return X, torch.stack(ys, dim=0), mask
