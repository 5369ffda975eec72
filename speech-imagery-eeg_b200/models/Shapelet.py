"""Shapelet bottleneck modules backed by the sm_100a kernels of libign_b200.so.

Same class names, constructor arguments, forward signatures, attributes and state_dict keys as the
reference's model/Shapelet.py (imported there as `models.Shapelet`), so exp/ and run.py drive them
unchanged; the eager unfold/broadcast/softmax chain of Shapelet.py:61-84 is replaced by one fused
kernel per length group (layers/shapelet_ops.py -> include/ign_b200.h).  CUDA-only.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from layers.shapelet_ops import (SeriesPack, instance_norm, resolve_dist, sbm_transform, shapelet_diversity,
                                 shapelet_transform)
from utils.shapelet_util import ModelInfo


class Shapelet(nn.Module):
    """K learned shapelets of one length per channel (reference Shapelet.py:46-87).

    forward(x[B,M,T]) -> (max_t RBF prob [B,K*M], min_t distance [B,K*M]), feature index k*M+m.
    `precision` is new: 'fp32' (CUDA-core), or a tcgen05 operand mode for the cross-term distances.
    """
    pool = "rbf_max"

    def __init__(self, dim_data, shapelet_len, num_shapelet=10, stride=1, eps=1., distance_func='euclidean',
                 memory_efficient=False, precision="fp32"):
        super().__init__()
        self.dim, self.length, self.n = int(dim_data), int(shapelet_len), int(num_shapelet)
        self.stride, self.eps = int(stride), eps
        self.distance_func, self.memory_efficient = distance_func, memory_efficient
        self.precision = precision
        # same initialiser (and RNG consumption) as Shapelet.py:57
        self.weights = nn.Parameter(torch.normal(0, 1, (self.n, self.dim, self.length)), requires_grad=True)

    def _dist(self):
        return resolve_dist(self.distance_func, self.memory_efficient)

    def _threshold(self):
        return None

    def transform(self, pack: SeriesPack):
        """Fused distance + pooling on a packed, normalised batch -> ([B,K*M], [B,K*M], argmin_t d [B,K,M])."""
        if pack.T < self.length:   # what x.unfold raises in Shapelet.py:61
            raise RuntimeError(f"maximum size for tensor at dimension 2 is {pack.T} but size is {self.length}")
        p, dmin, idx = shapelet_transform(pack, self.weights, self.stride, self.eps, self._dist(), self.pool,
                                          self._threshold(), self.precision)
        return p.flatten(start_dim=1), dmin.flatten(start_dim=1), idx

    def forward(self, x):
        p, dmin, _ = self.transform(SeriesPack.from_channel_major(x))
        return p, dmin

    def derivative(self):
        return torch.diff(self.weights, dim=-1)


class DistThresholdShapelet(Shapelet):
    """Learning-Time-series-Shapelets style layer (reference Shapelet.py:90-114): straight-through soft/hard
    min over time and p = sigmoid(threshold - min_d).  Like the reference it ignores cosine/pearson."""
    pool = "lts_min"

    def __init__(self, dim_data, shapelet_len, num_shapelet=10, stride=1, eps=1., distance_func='euclidean',
                 memory_efficient=False, precision="fp32"):
        super().__init__(dim_data, shapelet_len, num_shapelet, stride, eps, distance_func, memory_efficient, precision)
        self.threshold = nn.Parameter(torch.rand(1, self.n, self.dim).abs(), requires_grad=True)

    def _dist(self):
        return "sql2" if self.memory_efficient else "l1"   # Shapelet.py:100-103

    def _threshold(self):
        return self.threshold


class SelfAttention(nn.Module):
    """One-head attention over the shapelet features (reference Shapelet.py:117-131); tiny, stays PyTorch."""

    def __init__(self, dim_feature, dim_attn):
        super().__init__()
        self.q_proj = nn.Linear(1, dim_attn)
        self.k_proj = nn.Linear(1, dim_attn)
        self.pos_embed = nn.Embedding(num_embeddings=dim_feature, embedding_dim=dim_attn)

    def forward(self, x):
        pos = self.pos_embed(torch.arange(x.shape[1], device=x.device))
        v = x.unsqueeze(-1)
        return F.scaled_dot_product_attention(self.q_proj(v) + pos, self.k_proj(v) + pos, v).squeeze(-1)


def _group_length(frac, seq_len):
    return int(max(3, math.ceil(frac * seq_len)))          # Shapelet.py:153


def _group_stride(seq_len, length):
    return 1 if seq_len < 3000 else max(1, int(math.log2(length)))   # Shapelet.py:162


class ShapeBottleneckModel(nn.Module):
    """Instance norm -> one Shapelet per length fraction -> concat -> linear head (Shapelet.py:134-238)."""
    layer_cls = Shapelet

    def __init__(self, configs, num_shapelet=[5, 5, 5, 5], shapelet_len=[0.1, 0.2, 0.3, 0.5]):
        super().__init__()
        self.configs = configs
        self.num_shapelet = num_shapelet
        self.num_channel = configs.enc_in
        self.num_class = configs.num_class
        self.normalize = True
        self.shapelet_len = []
        self.shapelets = nn.ModuleList()
        self._build_layers(configs, num_shapelet, shapelet_len)
        self.total_shapelets = sum(num_shapelet * self.num_channel)   # list repetition, as Shapelet.py:167

        cls = configs.sbm_cls
        self.output_layer = nn.Linear(self.total_shapelets, self.num_class, bias=False)
        if cls == 'bilinear':
            self.output_bilinear = nn.Bilinear(self.total_shapelets, self.total_shapelets, self.num_class, bias=False)
        elif cls == 'attention':
            self.attention = SelfAttention(self.total_shapelets, 16)
        self.dropout = nn.Dropout(p=configs.dropout)
        self.lambda_reg = configs.lambda_reg      # L1 on classifier weights
        self.lambda_div = configs.lambda_div      # shapelet diversity
        self.last_indices = None                  # argmin_t d [B,K,M] per group of the latest forward
        self.loss_in_parent = False               # True when wrapped by InterpGN (it calls loss() itself)

    def _build_layers(self, configs, num_shapelet, shapelet_len):
        precision = getattr(configs, "shapelet_precision", "fp32")
        for i, frac in enumerate(shapelet_len):
            L = _group_length(frac, configs.seq_len)
            self.shapelets.append(self.layer_cls(
                dim_data=self.num_channel, shapelet_len=L, num_shapelet=num_shapelet[i], eps=configs.epsilon,
                distance_func=configs.distance_func, memory_efficient=configs.memory_efficient,
                stride=_group_stride(configs.seq_len, L), precision=precision))
            self.shapelet_len.append(L)

    def forward(self, x, *args, **kwargs):
        pack = instance_norm(x)                             # Shapelet.py:186-187 (one fused kernel)
        first = self.shapelets[0]                           # norm terms of all length groups in one pass
        pack.prepare_stats(first._dist(), [(s.length, s.stride) for s in self.shapelets])
        probs, dists, idxs = [], [], []
        for p, d, idx in sbm_transform(pack, self.shapelets):   # Shapelet.py:191-194, all groups as one autograd node
            probs.append(p.flatten(start_dim=1)); dists.append(d.flatten(start_dim=1)); idxs.append(idx)
        self.last_indices = idxs
        shapelet_probs = torch.cat(probs, dim=-1)
        shapelet_dists = torch.cat(dists, dim=-1)

        cls = self.configs.sbm_cls
        if cls == 'linear':
            out = self.output_layer(self.dropout(shapelet_probs))
        elif cls == 'bilinear':
            out = self.output_layer(self.dropout(shapelet_probs)) + \
                self.output_bilinear(self.dropout(shapelet_probs), self.dropout(shapelet_probs))
        elif cls == 'attention':
            out = self.output_layer(self.dropout(self.attention(shapelet_probs)))
        else:
            raise ValueError(f"unknown sbm_cls {cls!r}")
        # (InterpGN evaluates the regulariser itself, InterpGN.py:58; the reference computes it here as well and
        #  throws this copy away — the parent sets loss_in_parent to skip the duplicate)
        loss = None if self.loss_in_parent else self.loss().unsqueeze(0)
        return out, ModelInfo(d=shapelet_dists, p=shapelet_probs, shapelet_preds=out, preds=out, loss=loss)

    def step(self):
        with torch.no_grad():                               # Shapelet.py:212-215
            self.output_layer.weight.clamp_(0.)

    def loss(self):
        reg = self.output_layer.weight.abs().mean()
        div = self.diversity() if self.lambda_div > 0. else 0.
        return reg * self.lambda_reg + div * self.lambda_div

    def diversity(self):
        """mean_{m, i != j} exp(-||w_i - w_j + 1e-6||_2) per group (Shapelet.py:223-230; the 1e-6 is
        nn.PairwiseDistance's eps, added to the difference)."""
        total = 0.
        for layer in self.shapelets:
            if layer.weights.is_cuda:                                        # one fused launch each way per group
                total = total + shapelet_diversity(layer.weights)
                continue
            w = layer.weights.permute(1, 0, 2)                               # [M,K,L]  (host-side logic tests)
            dist = (w.unsqueeze(1) - w.unsqueeze(2) + 1e-6).norm(dim=-1)      # [M,K,K]
            off = 1.0 - torch.eye(w.shape[1], device=w.device, dtype=w.dtype)
            total = total + (torch.exp(-dist) * off).mean()
        return total

    def get_shapelets(self):
        out = []
        for layer in self.shapelets:
            w = layer.weights.data.cpu().numpy()
            out.extend((w[k, c, :], c) for k in range(w.shape[0]) for c in range(w.shape[1]))
        return out


class DistThresholdSBM(ShapeBottleneckModel):
    """SBM over DistThresholdShapelet layers (model 'LTS', reference Shapelet.py:241-262).  The reference
    appends to `shapelet_len` a second time; that harmless doubling is kept for attribute parity."""
    layer_cls = DistThresholdShapelet

    def _build_layers(self, configs, num_shapelet, shapelet_len):
        self.shapelet_len.extend(_group_length(f, configs.seq_len) for f in shapelet_len)
        super()._build_layers(configs, num_shapelet, shapelet_len)
