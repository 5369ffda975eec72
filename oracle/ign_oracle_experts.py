"""CPU restatement of the reference's deep experts (TEST INFRASTRUCTURE — not product code).

bench.py's `--impl reference` arm and `cpu_baseline` leg time the reference's eager CPU training step.  The shapelet
expert and gate come from oracle/ign_oracle.py; the deep expert that InterpGN mixes them with (InterpGN.py:41) is
restated here so that the reference arm imports nothing from the product package:

  FcnExpert          reference model/FullyConvNet.py:7-58
  TransformerExpert  reference model/Transformer.py:12-125 (classification branch), layers/Embed.py:8-126,
                     layers/Transformer_EncDec.py:27-80, layers/SelfAttention_Family.py:48-75,179-214
                     (einsum attention that materialises the B x H x T x T scores, as the reference does)

Both keep the reference's parameter names, so a reference state dict loads unchanged
(tests/test_oracle_experts.py checks outputs against the live modules when the reference tree is mounted).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class FcnExpert(nn.Module):
    """Three Conv1d + BatchNorm1d + ReLU blocks, global average pooling over time, one linear layer."""

    def __init__(self, enc_in, num_class, seq_len):
        super().__init__()
        widths = (3, 3, 2) if seq_len <= 10 else (8, 5, 3)          # FullyConvNet.py:11-47
        chans = (enc_in, 128, 256, 128)
        for i, k in enumerate(widths):
            setattr(self, "block%d" % (i + 1), nn.Sequential(nn.Conv1d(chans[i], chans[i + 1], k),
                                                           nn.BatchNorm1d(chans[i + 1]), nn.ReLU()))
        self.fc = nn.Linear(128, num_class)

    def forward(self, x, *unused):
        h = x.transpose(1, 2)                                       # 'b t c -> b c t'  (:52)
        h = self.block3(self.block2(self.block1(h)))
        return self.fc(h.mean(dim=-1))                              # AdaptiveAvgPool1d(1) + flatten (:56-57)


def _sinusoid_table(d_model, max_len=5000):
    pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    table = torch.zeros(max_len, d_model)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table.unsqueeze(0)


class _Named(nn.Module):
    """Empty container: gives nested parameter names (`encoder.attn_layers.0.attention.query_projection.weight`)."""


class TransformerExpert(nn.Module):
    def __init__(self, enc_in, num_class, seq_len, d_model=512, n_heads=8, d_ff=2048, e_layers=2, dropout=0.0,
                 activation="gelu"):
        super().__init__()
        self.n_heads, self.p_drop = n_heads, dropout
        self.ffn_act = F.relu if activation == "relu" else F.gelu
        emb = _Named()
        emb.value_embedding = _Named()
        emb.value_embedding.tokenConv = nn.Conv1d(enc_in, d_model, 3, padding=1, padding_mode="circular", bias=False)
        nn.init.kaiming_normal_(emb.value_embedding.tokenConv.weight, mode="fan_in", nonlinearity="leaky_relu")
        emb.position_embedding = _Named()
        emb.position_embedding.register_buffer("pe", _sinusoid_table(d_model))
        emb.temporal_embedding = _Named()                            # in the state dict, unused for classification
        emb.temporal_embedding.embed = nn.Linear(4, d_model, bias=False)
        self.enc_embedding = emb
        enc = _Named()
        enc.attn_layers = nn.ModuleList()
        for _ in range(e_layers):
            lay = _Named()
            lay.attention = _Named()
            for nm in ("query_projection", "key_projection", "value_projection", "out_projection"):
                setattr(lay.attention, nm, nn.Linear(d_model, d_model))
            lay.conv1 = nn.Conv1d(d_model, d_ff, 1)
            lay.conv2 = nn.Conv1d(d_ff, d_model, 1)
            lay.norm1, lay.norm2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
            enc.attn_layers.append(lay)
        enc.norm = nn.LayerNorm(d_model)
        self.encoder = enc
        self.projection = nn.Linear(d_model * seq_len, num_class)

    def _drop(self, t):
        return F.dropout(t, self.p_drop, self.training)

    def _attention(self, a, x):
        B, T, _ = x.shape
        H = self.n_heads
        q = a.query_projection(x).view(B, T, H, -1)
        k = a.key_projection(x).view(B, T, H, -1)
        v = a.value_projection(x).view(B, T, H, -1)
        scores = torch.einsum("blhe,bshe->bhls", q, k)               # SelfAttention_Family.py:61
        w = self._drop(torch.softmax(scores / math.sqrt(q.shape[-1]), dim=-1))
        return a.out_projection(torch.einsum("bhls,bshd->blhd", w, v).reshape(B, T, -1))

    def forward(self, x, x_mark_enc=None, *unused):
        e = self.enc_embedding
        h = e.value_embedding.tokenConv(x.permute(0, 2, 1)).transpose(1, 2) + e.position_embedding.pe[:, :x.shape[1]]
        h = self._drop(h)                                            # Embed.py:120-126 with x_mark None
        for lay in self.encoder.attn_layers:                        # Transformer_EncDec.py:39-51
            h = lay.norm1(h + self._drop(self._attention(lay.attention, h)))
            y = self._drop(self.ffn_act(lay.conv1(h.transpose(-1, 1))))
            y = self._drop(lay.conv2(y).transpose(-1, 1))
            h = lay.norm2(h + y)
        h = self._drop(F.gelu(self.encoder.norm(h)))                # Transformer.py:106-110
        if x_mark_enc is not None:
            h = h * x_mark_enc.unsqueeze(-1)
        return self.projection(h.reshape(h.shape[0], -1))


def build_expert(dnn_type, cfg):
    if dnn_type == "Transformer":
        return TransformerExpert(cfg.enc_in, cfg.num_class, cfg.seq_len, cfg.d_model, cfg.n_heads, cfg.d_ff,
                                 cfg.e_layers, cfg.dropout, cfg.activation)
    if dnn_type == "FCN":
        return FcnExpert(cfg.enc_in, cfg.num_class, cfg.seq_len)
    raise ValueError("oracle expert %r is not restated" % (dnn_type,))
