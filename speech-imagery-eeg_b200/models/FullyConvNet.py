"""FCN deep expert (3 x Conv1d+BN+ReLU -> global average pool -> Linear).  Out of the hot path: it stays
PyTorch/cuDNN, with the reference's module names and parameter shapes so state_dicts interchange
(reference model/FullyConvNet.py:7-58).

Layout: the batch arrives as [B,T,C] (time-major, channels contiguous) — exactly the NHWC image of a
[B,C,1,T] tensor.  On CUDA the convolutions therefore run as 1 x k conv2d in channels_last directly on that
view, which removes the NCHW<->NHWC transposes cuDNN otherwise inserts around every Conv1d (1.2 ms of a 27 ms
step at config 2).  Same arithmetic, same parameters; CPU tensors take the plain Conv1d path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_block(c_in, c_out, k):
    # BatchNorm2d and BatchNorm1d share parameter / buffer names and shapes; 2d accepts both [B,C,1,T] here
    return nn.Sequential(nn.Conv1d(c_in, c_out, k), nn.BatchNorm2d(c_out), nn.ReLU())


class FullyConvNetwork(nn.Module):
    def __init__(self, configs):
        super().__init__()
        kernels = (3, 3, 2) if configs.seq_len <= 10 else (8, 5, 3)
        self.block1 = _conv_block(configs.enc_in, 128, kernels[0])
        self.block2 = _conv_block(128, 256, kernels[1])
        self.block3 = _conv_block(256, 128, kernels[2])
        self.pooling = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Linear(128, configs.num_class)

    @staticmethod
    def _block(block, h):
        conv, bn, act = block[0], block[1], block[2]
        w = conv.weight.unsqueeze(2)                                  # [O,I,k] -> [O,I,1,k]
        if h.is_cuda:
            w = w.contiguous(memory_format=torch.channels_last)
        return act(bn(F.conv2d(h, w, conv.bias)))

    def forward(self, x, x_mark_enc=None, x_dec=None, x_mark_dec=None, mask=None):
        B, T, C = x.shape
        h = x.contiguous().view(B, 1, T, C).permute(0, 3, 1, 2)       # [B,C,1,T], channels_last strides, no copy
        h = self._block(self.block3, self._block(self.block2, self._block(self.block1, h)))
        return self.fc(h.mean(dim=(2, 3)))

