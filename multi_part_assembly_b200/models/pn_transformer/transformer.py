"""Inter-part attention: a pre-LN Transformer encoder over the P part tokens
with a key-padding mask (reference models/pn_transformer/transformer.py).
Parameters live in a stock `nn.TransformerEncoder` so reference checkpoints
load key for key; the forward runs the fused sm_100a path of `kernels`."""
import torch.nn as nn

from ... import kernels


def build_transformer_encoder(d_model, num_heads, ffn_dim, num_layers, norm_first=True,
                              dropout=0.1):
    layer = nn.TransformerEncoderLayer(
        d_model=d_model, nhead=num_heads, dim_feedforward=ffn_dim, dropout=dropout,
        norm_first=norm_first, batch_first=True)
    norm = nn.LayerNorm(d_model) if norm_first else None
    return nn.TransformerEncoder(encoder_layer=layer, num_layers=num_layers, norm=norm,
                                 enable_nested_tensor=False)


class TransformerEncoder(nn.Module):
    """tokens [B, P, C], valid_masks [B, P] (True = valid) -> [B, P, C]."""

    def __init__(self, d_model, num_heads, ffn_dim, num_layers, norm_first=True, dropout=0.1,
                 out_dim=None):
        super().__init__()
        self.transformer_encoder = build_transformer_encoder(
            d_model=d_model, num_heads=num_heads, ffn_dim=ffn_dim, num_layers=num_layers,
            norm_first=norm_first, dropout=dropout)
        self.out_fc = nn.Linear(d_model, out_dim) if out_dim is not None else nn.Identity()
        self.num_heads = num_heads
        self.dropout = dropout

    def forward(self, tokens, valid_masks):
        if valid_masks is not None:
            assert valid_masks.shape == tokens.shape[:2]
        out = kernels.transformer_forward(tokens, valid_masks, self.transformer_encoder,
                                          self.num_heads, self.training, self.dropout)
        return self.out_fc(out)
