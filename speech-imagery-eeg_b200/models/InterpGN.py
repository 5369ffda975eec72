"""Interpretability Gated Network: shapelet expert + deep expert mixed by a Gini gate
(reference model/InterpGN.py:22-66).  The gate/mixture is one fused kernel each way."""
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

    def forward(self, x, x_mark_enc=None, x_dec=None, x_mark_dec=None, mask=None, gating_value=None):
        sbm_out, info = self.sbm(x)
        deep_out = self.deep_model(x, x_mark_enc, x_dec, x_mark_dec, mask)
        # eta = (C*sum softmax(s)^2 - 1)/(C-1); out = eta*s + (1-eta)*z  (InterpGN.py:44-52)
        output, eta = gini_gate(sbm_out, deep_out, gating_value)
        return output, ModelInfo(d=info.d, p=info.p, eta=eta, shapelet_preds=sbm_out, dnn_preds=deep_out,
                                 preds=output, loss=self.loss().unsqueeze(0))

    def loss(self):
        return self.sbm.loss()

    def step(self):
        self.sbm.step()
