"""Drop-in mirror of reference model/attention.py (RotaryEmbedding, RelativeMultiHeadAttention): same
class names, constructor arguments, parameter/buffer names and shapes, computed by the
B200 kernels (fused QKV tcgen05 GEMM, RoPE, MQA flash attention)."""
import torch
import torch.nn as nn


class RotaryEmbedding(nn.Module):
    """reference model/attention.py:21-59.  Only `inv_freq` is persistent state; cos/sin tables are built on
    the device by the engine (fp32) and grown on demand."""

    def __init__(self, dim: int, max_seq_len: int = 5000, base: float = 10000.0):
        super().__init__()
        self.dim = dim
        self.max_seq_len = max_seq_len
        self.base = base
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, x: torch.Tensor, seq_len: int):
        t = torch.arange(seq_len, device=self.inv_freq.device, dtype=self.inv_freq.dtype)
        freqs = torch.outer(t, self.inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos()[None, None].to(x.dtype), emb.sin()[None, None].to(x.dtype)


class RelativeMultiHeadAttention(nn.Module):
    """reference model/attention.py:147-251 (use_mqa=True: one shared K/V head)."""

    def __init__(self, d_model: int, n_heads: int, dropout: float = 0.1, use_mqa: bool = True, use_flash: bool = True):
        super().__init__()
        assert d_model % n_heads == 0, "d_model must be divisible by n_heads"
        if not use_mqa:
            raise NotImplementedError("the B200 path implements the reference default use_mqa=True")
        self.d_model = d_model
        self.n_heads = n_heads
        self.d_head = d_model // n_heads
        self.use_mqa = use_mqa
        self.use_flash = use_flash
        self.n_kv_heads = 1
        self.rotary_emb = RotaryEmbedding(self.d_head)
        self.linear_q = nn.Linear(d_model, d_model)
        self.linear_k = nn.Linear(d_model, self.d_head)
        self.linear_v = nn.Linear(d_model, self.d_head)
        self.linear_out = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)  # unused, like the reference (model/attention.py:192)
        self.dropout_p = dropout

    def forward(self, x, x_k=None, x_v=None, mask=None):
        from ..functional import attention_module_forward
        for other in (x_k, x_v):
            if other is not None and other is not x and not (other.data_ptr() == x.data_ptr() and other.shape == x.shape):
                raise NotImplementedError("self-attention only (the reference always passes the same tensor three times)")
        return attention_module_forward(self, x, mask), None
