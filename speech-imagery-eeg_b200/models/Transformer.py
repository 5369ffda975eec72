"""Transformer deep expert for InterpGN (BASELINE config 5).  Out of the shapelet hot path: plain PyTorch, but
with the attention einsum pair of the reference (layers/SelfAttention_Family.py:48-75, which materialises
B*H*T^2 scores) replaced by scaled_dot_product_attention (flash kernels), SURVEY.md §8 row f2.

Module tree and parameter names follow the reference (model/Transformer.py:12-125, layers/Embed.py,
layers/Transformer_EncDec.py, layers/SelfAttention_Family.py:179-214) so state_dicts interchange:
  enc_embedding.{value_embedding.tokenConv, position_embedding.pe, temporal_embedding.embed}
  encoder.attn_layers.N.{attention.{query,key,value,out}_projection, conv1, conv2, norm1, norm2}, encoder.norm
  projection
Only the classification branch is built (the experiment passes task_name='classification').
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class PositionalEmbedding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pos = torch.arange(0, max_len).float().unsqueeze(1)
        div = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer('pe', pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, :x.size(1)]


class TokenEmbedding(nn.Module):
    def __init__(self, c_in, d_model):
        super().__init__()
        self.tokenConv = nn.Conv1d(c_in, d_model, kernel_size=3, padding=1, padding_mode='circular', bias=False)
        nn.init.kaiming_normal_(self.tokenConv.weight, mode='fan_in', nonlinearity='leaky_relu')

    def forward(self, x):
        return self.tokenConv(x.permute(0, 2, 1)).transpose(1, 2)


class TimeFeatureEmbedding(nn.Module):
    """Present in the reference's state_dict (embed='timeF'); unused by classification (x_mark is None)."""

    def __init__(self, d_model, freq='h'):
        super().__init__()
        d_inp = {'h': 4, 't': 5, 's': 6, 'm': 1, 'a': 1, 'w': 2, 'd': 3, 'b': 3}[freq]
        self.embed = nn.Linear(d_inp, d_model, bias=False)

    def forward(self, x):
        return self.embed(x)


class DataEmbedding(nn.Module):
    def __init__(self, c_in, d_model, embed_type='timeF', freq='h', dropout=0.1):
        super().__init__()
        self.value_embedding = TokenEmbedding(c_in, d_model)
        self.position_embedding = PositionalEmbedding(d_model)
        self.temporal_embedding = TimeFeatureEmbedding(d_model, freq)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x, x_mark=None):
        out = self.value_embedding(x) + self.position_embedding(x)
        if x_mark is not None:
            out = out + self.temporal_embedding(x_mark)
        return self.dropout(out)


class AttentionLayer(nn.Module):
    def __init__(self, d_model, n_heads, attn_dropout=0.0):
        super().__init__()
        self.n_heads = n_heads
        self.attn_dropout = attn_dropout
        self.query_projection = nn.Linear(d_model, d_model)
        self.key_projection = nn.Linear(d_model, d_model)
        self.value_projection = nn.Linear(d_model, d_model)
        self.out_projection = nn.Linear(d_model, d_model)

    def forward(self, x):
        B, L, D = x.shape
        H = self.n_heads
        q = self.query_projection(x).view(B, L, H, -1).transpose(1, 2)
        k = self.key_projection(x).view(B, L, H, -1).transpose(1, 2)
        v = self.value_projection(x).view(B, L, H, -1).transpose(1, 2)
        # softmax(q k^T / sqrt(E)) v, un-masked (FullAttention(mask_flag=False)), without materialising B*H*L^2
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=self.attn_dropout if self.training else 0.0)
        return self.out_projection(o.transpose(1, 2).reshape(B, L, D))


class EncoderLayer(nn.Module):
    def __init__(self, d_model, n_heads, d_ff, dropout, activation):
        super().__init__()
        self.attention = AttentionLayer(d_model, n_heads, dropout)
        self.conv1 = nn.Conv1d(d_model, d_ff, kernel_size=1)
        self.conv2 = nn.Conv1d(d_ff, d_model, kernel_size=1)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.activation = F.relu if activation == "relu" else F.gelu

    def forward(self, x):
        x = self.norm1(x + self.dropout(self.attention(x)))
        # the two kernel-size-1 convolutions are per-position linears: run them as GEMMs on [B,L,D]
        y = self.dropout(self.activation(F.linear(x, self.conv1.weight.squeeze(-1), self.conv1.bias)))
        y = self.dropout(F.linear(y, self.conv2.weight.squeeze(-1), self.conv2.bias))
        return self.norm2(x + y)


class Encoder(nn.Module):
    def __init__(self, layers, norm_layer):
        super().__init__()
        self.attn_layers = nn.ModuleList(layers)
        self.norm = norm_layer

    def forward(self, x):
        for layer in self.attn_layers:
            x = layer(x)
        return self.norm(x)


class Model(nn.Module):
    def __init__(self, configs):
        super().__init__()
        self.task_name = getattr(configs, "task_name", "classification")
        if self.task_name != 'classification':
            raise ValueError("only the classification branch of the Transformer expert is built here")
        self.enc_embedding = DataEmbedding(configs.enc_in, configs.d_model, configs.embed, configs.freq, configs.dropout)
        self.encoder = Encoder([EncoderLayer(configs.d_model, configs.n_heads, configs.d_ff, configs.dropout,
                                             configs.activation) for _ in range(configs.e_layers)],
                               norm_layer=nn.LayerNorm(configs.d_model))
        self.act = F.gelu
        self.dropout = nn.Dropout(configs.dropout)
        self.projection = nn.Linear(configs.d_model * configs.seq_len, configs.num_class)

    def forward(self, x_enc, x_mark_enc, x_dec=None, x_mark_dec=None, mask=None):
        out = self.encoder(self.enc_embedding(x_enc, None))
        out = self.dropout(self.act(out))
        if x_mark_enc is not None:
            out = out * x_mark_enc.unsqueeze(-1)        # zero the padded positions (reference :111)
        return self.projection(out.reshape(out.shape[0], -1))
