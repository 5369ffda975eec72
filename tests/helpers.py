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


def seeded_fill(model, seed):
    """Overwrite every floating-point parameter of `model` with values drawn from a CPU generator seeded by (seed, name):
    N(0,1) for the shapelets (their initialiser, Shapelet.py:57), N(0, 1/fan_in) for matrices / convolutions, small
    values for biases and unit-ish scales for 1-D weights.  The full-size golden cases store no state dict (4-100 MB);
    the golden generator (live reference) and the tests (this repo's modules, same state-dict keys) both call this
    instead, and the fixture carries a checksum of what it produced."""
    import zlib
    digest = 0.0
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            g = torch.Generator().manual_seed(int(seed) * 1000003 + zlib.crc32(name.encode()) % 1000003)
            v = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if name.endswith("weights") or name.endswith("threshold"):
                v = v.abs() if name.endswith("threshold") else v
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                v = v / float(np.sqrt(max(1, fan_in)))
            elif name.endswith("bias"):
                v = 0.05 * v
            else:
                v = 1.0 + 0.1 * v
            p.copy_(v.to(p.dtype))
            digest += float(v.double().abs().sum())
    return digest


def seeded_batch(B, T, M, C, seed):
    """(x [B,T,M], y [B]) from a CPU generator — regenerated identically by the golden generator and the tests."""
    g = torch.Generator().manual_seed(int(seed))
    x = torch.randn(B, T, M, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    return x, y


def sample_indices(numel, n, seed):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, numel, (min(n, numel),), generator=g)
