"""Boundary containers of the shapelet models — same field names as the reference's
utils/shapelet_util.py:17-41 (ModelInfo is the second return value of every SBM-family forward;
ClassificationResult is what Experiment.test() hands back).  The reference file's plotting helpers
(seaborn / matplotlib / t-SNE, :44-194) are out of scope for the hot path and are not reproduced."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class ModelInfo:
    d: Optional[torch.Tensor] = None               # [B,F] min window distance per shapelet feature
    p: Optional[torch.Tensor] = None               # [B,F] pooled predicate
    eta: Optional[torch.Tensor] = None             # [B,1] Gini gate (InterpGN only)
    shapelet_preds: Optional[torch.Tensor] = None  # [B,C] shapelet-expert logits
    dnn_preds: Optional[torch.Tensor] = None       # [B,C] deep-expert logits (InterpGN only)
    preds: Optional[torch.Tensor] = None           # [B,C] final logits
    loss: Optional[torch.Tensor] = None            # [1]   regulariser


@dataclass
class ClassificationResult:
    x_data: Optional[torch.Tensor] = None
    shapelets: Optional[object] = None
    trues: Optional[torch.Tensor] = None
    preds: Optional[torch.Tensor] = None
    shapelet_preds: Optional[torch.Tensor] = None
    dnn_preds: Optional[torch.Tensor] = None
    p: Optional[torch.Tensor] = None
    d: Optional[torch.Tensor] = None
    w: Optional[torch.Tensor] = None
    eta: Optional[torch.Tensor] = None
    loss: Optional[float] = None
    accuracy: Optional[float] = None
