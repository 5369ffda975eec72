"""CPU oracle for the InterpGN shapelet hot path (TEST INFRASTRUCTURE — not product code).

This file is a plain, chunked, CPU-only torch restatement of the reference algorithm in
/root/reference/InterpretGatedNetwork/model/Shapelet.py and model/InterpGN.py.  It is the
checker the CUDA path is compared with.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product package never does (the
product fails loudly without its CUDA library).

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the authoring container by
tests/golden/make_golden.py (which imports the unmodified reference through oracle/ref_shim.py)
and committed as tests/golden/*.npz; tests/test_oracle_golden.py replays them.

Every function cites the reference lines it restates.  Shapes follow the reference:
  x  [B, T, M]   raw batch (time-major, channels last)
  xn [B, M, T]   instance-normalised, channel-major
  W  [K, M, L]   K shapelets of length L per channel (Shapelet.weights)
  d  [B, T', K, M] window distances,  T' = (T - L) // stride + 1
  features are flattened as k * M + m  (Shapelet.py:84)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

DIST_L1 = "euclidean"       # Shapelet.py:74   mean |x - w|   (the reference calls this 'euclidean')
DIST_SQL2 = "sql2"          # Shapelet.py:28   mean (w - x)^2 (memory_efficient arithmetic)
DIST_COS = "cosine"         # Shapelet.py:64-66
DIST_PEARSON = "pearson"    # Shapelet.py:11-19, 67-69

POOL_RBF_MAX = "rbf_max"    # Shapelet.py:77-84       (class Shapelet)
POOL_LTS_MIN = "lts_min"    # Shapelet.py:105-111     (class DistThresholdShapelet)


def resolve_mode(distance_func: str, memory_efficient: bool) -> str:
    """Flag -> arithmetic, as dispatched by Shapelet.forward (Shapelet.py:64-74)."""
    if distance_func == "cosine":
        return DIST_COS
    if distance_func == "pearson":
        return DIST_PEARSON
    return DIST_SQL2 if memory_efficient else DIST_L1


def shapelet_lengths(seq_len: int, fracs: Sequence[float]) -> List[int]:
    """Shapelet.py:153  L = max(3, ceil(frac * seq_len))."""
    return [int(max(3, math.ceil(f * seq_len))) for f in fracs]


def shapelet_stride(seq_len: int, length: int) -> int:
    """Shapelet.py:162  stride = 1 if seq_len < 3000 else max(1, int(log2 L))."""
    return 1 if seq_len < 3000 else max(1, int(math.log2(length)))


def num_windows(T: int, L: int, stride: int) -> int:
    """x.unfold(2, L, stride) (Shapelet.py:61); raises like unfold when T < L."""
    if T < L:
        raise RuntimeError("maximum size for tensor at dimension 2 is %d but size is %d" % (T, L))
    return (T - L) // stride + 1


def instance_norm(x: torch.Tensor) -> torch.Tensor:
    """Shapelet.py:186-187: 'b t c -> b c t', (x - mean_T) / (std_T(unbiased) + 1e-8)."""
    xc = x.transpose(1, 2)
    return (xc - xc.mean(dim=-1, keepdim=True)) / (xc.std(dim=-1, keepdim=True) + 1e-8)


def window_distance(xn: torch.Tensor, W: torch.Tensor, stride: int, mode: str) -> torch.Tensor:
    """d[B,T',K,M] for one length group — Shapelet.py:61-74 (and :28 for sql2).

    Chunked over the batch so the 5-D temporary [1,T',K,M,L] of Shapelet.py:74 stays bounded.
    """
    B, M, T = xn.shape
    K, M2, L = W.shape
    assert M == M2
    num_windows(T, L, stride)
    out = []
    for b in range(B):
        xw = xn[b:b + 1].unfold(2, L, stride)             # [1,M,T',L]      Shapelet.py:61
        xw = xw.permute(0, 2, 1, 3).unsqueeze(2)           # [1,T',1,M,L]    Shapelet.py:62
        if mode == DIST_COS:                               # Shapelet.py:64-66
            d = 1.0 - torch.nn.functional.cosine_similarity(xw, W, dim=-1)
        elif mode == DIST_PEARSON:                         # Shapelet.py:11-19,67-69
            xc = xw - xw.mean(dim=-1, keepdim=True)
            wc = W - W.mean(dim=-1, keepdim=True)
            num = (xc * wc).sum(dim=-1)
            den = torch.sqrt((xc ** 2).sum(dim=-1) * (wc ** 2).sum(dim=-1)) + 1e-8
            d = 1.0 - num / den
        elif mode == DIST_SQL2:                            # Shapelet.py:28
            d = (W - xw).pow(2).mean(dim=-1)
        elif mode == DIST_L1:                              # Shapelet.py:74
            d = (xw - W).abs().mean(dim=-1)
        else:
            raise ValueError(mode)
        out.append(d)
    return torch.cat(out, dim=0)


def rbf(d: torch.Tensor, eps: float) -> torch.Tensor:
    """Shapelet.py:77  p = exp(-(eps*d)^2)."""
    return torch.exp(-torch.pow(eps * d, 2))


def ste_max_pool(p: torch.Tensor) -> torch.Tensor:
    """Shapelet.py:79-82 straight-through soft/hard max over time (dim 1)."""
    hard = torch.zeros_like(p).scatter_(1, p.argmax(dim=1, keepdim=True), 1.0)
    soft = torch.softmax(p, dim=1)
    onehot = hard + soft - soft.detach()
    return torch.sum(onehot * p, dim=1)


def ste_min_pool(d: torch.Tensor) -> torch.Tensor:
    """Shapelet.py:105-108 straight-through soft/hard min over time (dim 1)."""
    hard = torch.zeros_like(d).scatter_(1, d.argmin(dim=1, keepdim=True), 1.0)
    soft = torch.nn.functional.softmin(d, dim=1)
    onehot = hard + soft - soft.detach()
    return torch.sum(onehot * d, dim=1)


@dataclass
class ShapeletOut:
    p: torch.Tensor            # [B, K*M]  pooled predicate (max RBF prob, or sigmoid(thr - min d))
    dmin: torch.Tensor         # [B, K*M]  min_t d
    arg_hard: torch.Tensor     # [B, K, M] index the STE's hard one-hot selects (argmax_t p / argmin_t d)
    argmin_d: torch.Tensor     # [B, K, M] argmin_t d (first index)
    d: torch.Tensor            # [B, T', K, M] (kept for tests)


def shapelet_forward(xn, W, stride=1, eps=1.0, mode=DIST_L1, pool=POOL_RBF_MAX,
                     threshold: Optional[torch.Tensor] = None) -> ShapeletOut:
    """Shapelet.forward (Shapelet.py:60-84) / DistThresholdShapelet.forward (:96-111)."""
    d = window_distance(xn, W, stride, mode)
    if pool == POOL_RBF_MAX:
        p = rbf(d, eps)
        pooled = ste_max_pool(p)
        arg_hard = p.argmax(dim=1)
    else:
        min_d = ste_min_pool(d)
        pooled = torch.sigmoid(threshold - min_d)          # Shapelet.py:109, threshold [1,K,M]
        arg_hard = d.argmin(dim=1)
    return ShapeletOut(p=pooled.flatten(start_dim=1), dmin=d.min(dim=1).values.flatten(start_dim=1),
                       arg_hard=arg_hard, argmin_d=d.argmin(dim=1), d=d)


# ----------------------------------------------------------------------------------------------
# Explicit backward restatement (what autograd derives from Shapelet.py:74-82).
# The CUDA backward implements these closed forms; tests check them against autograd here.
# ----------------------------------------------------------------------------------------------

def pooled_grad_wrt_d(d, g, eps, pool, threshold=None):
    """c[B,T',K,M] = dLoss/dd given g[B,K,M] = dLoss/d(pooled).

    rbf_max: dp_max/dp_t = hard_t + soft_t (p_t - pbar)      (Shapelet.py:79-82)
             dp_t/dd_t   = p_t * (-2 eps^2 d_t)               (Shapelet.py:77)
    lts_min: dmin/dd_t   = hard_t - soft_t (d_t - dbar),  soft = softmin(d)   (Shapelet.py:105-108)
             dp/dmin     = -p (1 - p),  p = sigmoid(thr - min_d)               (Shapelet.py:109)
    """
    if pool == POOL_RBF_MAX:
        p = rbf(d, eps)
        hard = torch.zeros_like(p).scatter_(1, p.argmax(dim=1, keepdim=True), 1.0)
        soft = torch.softmax(p, dim=1)
        pbar = (soft * p).sum(dim=1, keepdim=True)
        dp = g.unsqueeze(1) * (hard + soft * (p - pbar))
        return dp * p * (-2.0 * eps * eps * d)
    hard = torch.zeros_like(d).scatter_(1, d.argmin(dim=1, keepdim=True), 1.0)
    soft = torch.nn.functional.softmin(d, dim=1)
    dbar = (soft * d).sum(dim=1, keepdim=True)
    dmin = d.min(dim=1, keepdim=True).values
    pk = torch.sigmoid(threshold.unsqueeze(1) - dmin)      # [B,1,K,M]
    gm = g.unsqueeze(1) * (-pk * (1.0 - pk))
    return gm * (hard - soft * (d - dbar))


def weight_grad_from_c(xn, W, c, stride, mode):
    """dW[K,M,L] = sum_{b,t} c[b,t,k,m] * dd_t/dW[k,m,:]  (closed forms)."""
    B, M, T = xn.shape
    K, _, L = W.shape
    dW = torch.zeros_like(W)
    for b in range(B):
        xw = xn[b].unfold(1, L, stride)                     # [M,T',L]
        cb = c[b]                                           # [T',K,M]
        if mode == DIST_L1:
            # dd/dw_l = -sign(x_{t+l} - w_l)/L, sign(0)=0
            s = torch.sign(xw.permute(1, 0, 2).unsqueeze(1) - W.unsqueeze(0))     # [T',K,M,L]
            dW += -(cb.unsqueeze(-1) * s).sum(dim=0) / L
        elif mode == DIST_SQL2:
            # dd/dw_l = 2 (w_l - x_{t+l}) / L                                     (Shapelet.py:34-39)
            G = torch.einsum("tkm,mtl->kml", cb, xw)
            dW += (2.0 / L) * (W * cb.sum(dim=0).unsqueeze(-1) - G)
        elif mode == DIST_COS:
            nx = xw.norm(dim=-1).clamp_min(1e-8)            # [M,T']
            nw = W.norm(dim=-1).clamp_min(1e-8)             # [K,M]
            cosv = torch.einsum("mtl,kml->tkm", xw, W) / (nx.t().unsqueeze(1) * nw.unsqueeze(0))
            a = cb / nx.t().unsqueeze(1)                    # [T',K,M]
            G = torch.einsum("tkm,mtl->kml", a, xw)
            dW += -(G / nw.unsqueeze(-1)) + W * ((cb * cosv).sum(dim=0) / (nw * nw)).unsqueeze(-1)
        elif mode == DIST_PEARSON:
            mu = xw.mean(dim=-1)                            # [M,T']
            wc = W - W.mean(dim=-1, keepdim=True)
            Sx = ((xw - mu.unsqueeze(-1)) ** 2).sum(dim=-1)  # [M,T']
            Sw = (wc ** 2).sum(dim=-1)                       # [K,M]
            root = torch.sqrt(Sx.t().unsqueeze(1) * Sw.unsqueeze(0))              # [T',K,M]
            D = root + 1e-8
            num = torch.einsum("mtl,kml->tkm", xw - mu.unsqueeze(-1), wc)
            corr = num / D
            a = cb / D
            G = torch.einsum("tkm,mtl->kml", a, xw)
            amu = (a * mu.t().unsqueeze(1)).sum(dim=0)      # [K,M]
            # dD/dwc_l = Sx * wc_l / root
            coef = (cb * corr * Sx.t().unsqueeze(1) / (root * D)).sum(dim=0)       # [K,M]
            dW += -(G - amu.unsqueeze(-1)) + wc * coef.unsqueeze(-1)
        else:
            raise ValueError(mode)
    return dW


def shapelet_backward_formula(xn, W, g, stride=1, eps=1.0, mode=DIST_L1, pool=POOL_RBF_MAX,
                              threshold=None):
    """Closed-form dW (and dthreshold for lts_min) for upstream g[B,K,M] on the pooled output."""
    d = window_distance(xn, W, stride, mode)
    c = pooled_grad_wrt_d(d, g, eps, pool, threshold)
    dW = weight_grad_from_c(xn, W, c, stride, mode)
    dthr = None
    if pool == POOL_LTS_MIN:
        dmin = d.min(dim=1).values
        pk = torch.sigmoid(threshold - dmin)
        dthr = (g * pk * (1.0 - pk)).sum(dim=0, keepdim=True)
    return dW, dthr


def shapelet_backward_autograd(xn, W, g, stride=1, eps=1.0, mode=DIST_L1, pool=POOL_RBF_MAX,
                               threshold=None, need_dx=False):
    """The same gradients obtained by autograd through the restated forward (ground truth)."""
    Wv = W.detach().clone().requires_grad_(True)
    xv = xn.detach().clone().requires_grad_(need_dx)
    tv = threshold.detach().clone().requires_grad_(True) if threshold is not None else None
    out = shapelet_forward(xv, Wv, stride, eps, mode, pool, tv)
    K, M = W.shape[0], W.shape[1]
    (out.p.view(-1, K, M) * g).sum().backward()
    return Wv.grad, (tv.grad if tv is not None else None), (xv.grad if need_dx else None)


# ----------------------------------------------------------------------------------------------
# Bottleneck model, regulariser and gate
# ----------------------------------------------------------------------------------------------

def diversity(weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """ShapeBottleneckModel.diversity (Shapelet.py:223-230); nn.PairwiseDistance adds eps=1e-6
    to the difference before the 2-norm."""
    loss = 0.0
    for Wg in weights:
        sh = Wg.permute(1, 0, 2)                                        # [M,K,L]
        diff = sh.unsqueeze(1) - sh.unsqueeze(2) + 1e-6                 # [M,K,K,L]
        dist = diff.norm(dim=-1)
        mask = torch.ones_like(dist) - torch.eye(sh.shape[1]).unsqueeze(0).to(dist)
        loss = loss + (torch.exp(-dist) * mask).mean()
    return loss


def sbm_loss(Wc: torch.Tensor, weights: Sequence[torch.Tensor], lambda_reg: float, lambda_div: float):
    """ShapeBottleneckModel.loss (Shapelet.py:217-221)."""
    loss_reg = Wc.abs().mean()
    loss_div = diversity(weights) if lambda_div > 0.0 else 0.0
    return loss_reg * lambda_reg + loss_div * lambda_div


def sbm_forward(x, weights: Sequence[torch.Tensor], strides: Sequence[int], Wc: torch.Tensor,
                eps=1.0, mode=DIST_L1, pool=POOL_RBF_MAX, thresholds=None):
    """ShapeBottleneckModel.forward with sbm_cls='linear', dropout 0 (Shapelet.py:184-210).
    Returns (logits[B,C], probs[B,F], dists[B,F])."""
    xn = instance_norm(x)
    ps, ds = [], []
    for i, Wg in enumerate(weights):
        o = shapelet_forward(xn, Wg, strides[i], eps, mode, pool,
                             None if thresholds is None else thresholds[i])
        ps.append(o.p)
        ds.append(o.dmin)
    probs = torch.cat(ps, dim=-1)
    dists = torch.cat(ds, dim=-1)
    return probs @ Wc.t(), probs, dists


def gate_forward(sbm_out: torch.Tensor, deep_out: torch.Tensor, gating_value: Optional[float] = None):
    """InterpGN.forward gate + mixture (InterpGN.py:44-52). Returns (out[B,C], eta[B,1])."""
    q = torch.softmax(sbm_out, dim=-1)
    c = sbm_out.shape[-1]
    gini = q.pow(2).sum(-1, keepdim=True)
    eta = (c * gini - 1) / (c - 1)
    if gating_value is not None:
        mask = (eta > gating_value).float()
        eta = torch.ones_like(eta) * mask + eta * (1 - mask)
    out = eta * sbm_out + (1.0 - eta) * deep_out
    return out, eta


def gate_backward_formula(sbm_out, deep_out, g_out, g_eta=None, gating_value=None):
    """Closed-form gradients of the gate (eta is NOT detached in InterpGN.py:47-52).

    out = eta*s + (1-eta)*z;  eta = (C*sum q^2 - 1)/(C-1), q = softmax(s)
    d eta / d s_j = (2C/(C-1)) * q_j * (q_j - sum_i q_i^2)
    Where the hard gate fires (eta > gating_value) eta is the constant 1 (mask has no grad).
    g_eta is an optional upstream gradient on the returned eta itself.
    """
    q = torch.softmax(sbm_out, dim=-1)
    C = sbm_out.shape[-1]
    gini = q.pow(2).sum(-1, keepdim=True)
    eta_raw = (C * gini - 1) / (C - 1)
    live = torch.ones_like(eta_raw)
    eta = eta_raw
    if gating_value is not None:
        fired = (eta_raw > gating_value).to(eta_raw)
        live = 1.0 - fired
        eta = fired + eta_raw * live
    d_eta = (g_out * (sbm_out - deep_out)).sum(-1, keepdim=True)
    if g_eta is not None:
        d_eta = d_eta + g_eta
    d_eta = d_eta * live
    deta_ds = (2.0 * C / (C - 1)) * q * (q - gini)
    g_s = eta * g_out + d_eta * deta_ds
    g_z = (1.0 - eta) * g_out
    return g_s, g_z
