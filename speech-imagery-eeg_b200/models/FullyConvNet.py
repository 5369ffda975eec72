"""FCN deep expert (3 x Conv1d+BN+ReLU -> global average pool -> Linear).  Out of the hot path: it stays
plain PyTorch/cuDNN, with the reference's module names so state_dicts interchange
(reference model/FullyConvNet.py:7-58)."""
import torch.nn as nn


def _conv_block(c_in, c_out, k):
    return nn.Sequential(nn.Conv1d(c_in, c_out, k), nn.BatchNorm1d(c_out), nn.ReLU())


class FullyConvNetwork(nn.Module):
    def __init__(self, configs):
        super().__init__()
        kernels = (3, 3, 2) if configs.seq_len <= 10 else (8, 5, 3)
        self.block1 = _conv_block(configs.enc_in, 128, kernels[0])
        self.block2 = _conv_block(128, 256, kernels[1])
        self.block3 = _conv_block(256, 128, kernels[2])
        self.pooling = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Linear(128, configs.num_class)

    def forward(self, x, x_mark_enc=None, x_dec=None, x_mark_dec=None, mask=None):
        h = x.transpose(1, 2)                       # [B,T,C] -> [B,C,T]
        h = self.block3(self.block2(self.block1(h)))
        return self.fc(self.pooling(h).flatten(start_dim=1))
