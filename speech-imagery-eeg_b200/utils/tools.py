"""Training utilities with the reference's semantics (utils/tools.py:9-77): early stopping that
checkpoints the best state_dict, ETA formatting, and the Gini coefficient of classifier weights."""
import os

import numpy as np
import torch


class EarlyStopping:
    """Tracks -score; saves `checkpoint.pth` on every improvement (reference utils/tools.py:9-38,
    with np.Inf -> np.inf so it runs on numpy >= 2)."""

    def __init__(self, patience=7, verbose=False, delta=0, is_main=True):
        self.patience, self.verbose, self.delta = patience, verbose, delta
        self.counter, self.best_score, self.early_stop = 0, None, False
        self.val_loss_min = np.inf
        self.is_main = is_main          # under data-parallel training only rank 0 writes

    def __call__(self, val_loss, model, path):
        score = -val_loss
        if self.best_score is None or score >= self.best_score + self.delta:
            self.best_score = score
            self.save_checkpoint(val_loss, model, path)
            self.counter = 0
        else:
            self.counter += 1
            if self.is_main:
                print(f"EarlyStopping counter: {self.counter} out of {self.patience}")
            self.early_stop = self.counter >= self.patience

    def save_checkpoint(self, val_loss, model, path):
        if self.is_main:
            if self.verbose:
                print(f"Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f}).  Saving model ...")
            os.makedirs(path, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(path, "checkpoint.pth"))
        self.val_loss_min = val_loss


def convert_to_hms(seconds):
    s = int(seconds)
    return f"{s // 3600:02d}:{(s % 3600) // 60:02d}:{s % 60:02d}"


def gini_coefficient(w):
    """Mean Gini coefficient over the rows of a non-negative matrix (reference utils/tools.py:54-77)."""
    w = np.asarray(w, dtype=np.float64)
    if w.shape[1] == 0:
        return 0.0
    srt = np.sort(w, axis=1)
    n = w.shape[1]
    idx = np.arange(1, n + 1)
    g = (2.0 * (srt * idx).sum(axis=1)) / (n * srt.sum(axis=1)) - (n + 1) / n
    return float(np.mean(g))
