"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

import ign_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# reference flag -> (oracle mode, kernel dist name)
MODES = {"euclidean": (O.DIST_L1, "l1"), "sql2": (O.DIST_SQL2, "sql2"),
         "cosine": (O.DIST_COS, "cosine"), "pearson": (O.DIST_PEARSON, "pearson")}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def t(a, device="cpu", dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(device=device, dtype=dtype)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol, atol, what=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    bad = err > bound
    assert not bool(bad.any()), "%s: %d/%d elements out of tolerance (rtol %g atol %g); max abs err %g, max |ref| %g" % (
        what, int(bad.sum()), bad.numel(), rtol, atol, float(err.max()), float(b.abs().max()))


def index_parity(idx_ours, idx_ref, score_ref, larger_is_better, L, what=""):
    """Indices must match, except where the reference's own scores of the two candidates are within
    fp32 summation-order noise: |score_ref[ours] - score_ref[ref]| <= 4 ulp*sqrt(L)."""
    idx_ours = idx_ours.cpu().long()
    idx_ref = idx_ref.cpu().long()
    mism = idx_ours != idx_ref
    n_mis = int(mism.sum())
    if n_mis == 0:
        return 0
    # score_ref: [B,T',K,M]; gather along dim 1
    so = score_ref.gather(1, idx_ours.unsqueeze(1)).squeeze(1)
    sr = score_ref.gather(1, idx_ref.unsqueeze(1)).squeeze(1)
    tol = 4.0 * np.finfo(np.float32).eps * np.sqrt(L) * sr.abs().clamp_min(1e-30)
    really_bad = mism & ((so - sr).abs() > tol)
    assert not bool(really_bad.any()), "%s: %d index mismatches beyond near-tie tolerance" % (what, int(really_bad.sum()))
    return n_mis
