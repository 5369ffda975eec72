"""Interpretability Gated Network: shapelet expert + deep expert mixed by a Gini gate
(reference model/InterpGN.py:22-66).  The gate/mixture is one fused kernel each way.

The two experts are independent until the gate, and they stress different parts of the SM: the default (L1) shapelet
kernels are bound by the FP32 issue slots, the deep expert's cuDNN / ATen kernels by tensor cores and HBM (batch norm,
bias and reduction passes).  On CUDA the deep expert therefore runs on a side stream, forward and — because autograd
replays every node on its forward stream — backward, so its memory-bound passes hide under the shapelet kernels."""
import os

import torch
import torch.nn as nn

from layers.shapelet_ops import gini_gate
from models.FullyConvNet import FullyConvNetwork
from models.Shapelet import ShapeBottleneckModel
from utils.shapelet_util import ModelInfo


def _transformer(configs):
    from models.Transformer import Model
    return Model(configs)


# PatchTST / TimesNet / ResNet experts of the reference are outside the hot path and not built here.
dnn_dict = {'FCN': FullyConvNetwork, 'Transformer': _transformer}


class InterpGN(nn.Module):
    def __init__(self, configs, num_shapelet=[5, 5, 5, 5], shapelet_len=[0.1, 0.2, 0.3, 0.5]):
        super().__init__()
        self.configs = configs
        self.sbm = ShapeBottleneckModel(configs=configs, num_shapelet=num_shapelet, shapelet_len=shapelet_len)
        self.sbm.loss_in_parent = True             # forward() below evaluates the regulariser once
        if configs.dnn_type not in dnn_dict:
            raise ValueError(f"dnn_type {configs.dnn_type!r} is not built in this framework "
                             f"(available: {sorted(dnn_dict)})")
        self.deep_model = dnn_dict[configs.dnn_type](configs)
        self.overlap_experts = os.environ.get("IGN_OVERLAP_EXPERTS", "1") != "0"
        self._side = None

    def _experts(self, x, x_mark_enc, x_dec, x_mark_dec, mask):
        if not (self.overlap_experts and x.is_cuda):
            sbm_out, info = self.sbm(x)
            return sbm_out, info, self.deep_model(x, x_mark_enc, x_dec, x_mark_dec, mask)
        cur = torch.cuda.current_stream(x.device)
        if self._side is None or self._side.device != x.device:
            # default priority: a high-priority side stream was measured slower (22.85 vs 22.55 ms per step at config 2)
            self._side = torch.cuda.Stream(device=x.device)
        side = self._side
        side.wait_stream(cur)                               # x (and the parameters) are ready
        with torch.cuda.stream(side):
            deep_out = self.deep_model(x, x_mark_enc, x_dec, x_mark_dec, mask)
        for t in (x, x_mark_enc, mask):                      # the caching allocator must not recycle them under the side stream
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(side)
        sbm_out, info = self.sbm(x)
        cur.wait_stream(side)
        deep_out.record_stream(cur)
        return sbm_out, info, deep_out

    def forward(self, x, x_mark_enc=None, x_dec=None, x_mark_dec=None, mask=None, gating_value=None):
        sbm_out, info, deep_out = self._experts(x, x_mark_enc, x_dec, x_mark_dec, mask)
        # eta = (C*sum softmax(s)^2 - 1)/(C-1); out = eta*s + (1-eta)*z  (InterpGN.py:44-52)
        output, eta = gini_gate(sbm_out, deep_out, gating_value)
        return output, ModelInfo(d=info.d, p=info.p, eta=eta, shapelet_preds=sbm_out, dnn_preds=deep_out,
                                 preds=output, loss=self.loss().unsqueeze(0))

    def loss(self):
        return self.sbm.loss()

    def step(self):
        self.sbm.step()
