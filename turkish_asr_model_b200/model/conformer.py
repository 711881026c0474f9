"""Drop-in mirror of reference model/conformer.py: same class names, constructor signatures and
state_dict keys/shapes (incl. the dead `norm_conv` parameters and BatchNorm buffers), so
reference checkpoints load with load_state_dict and vice versa.  The arithmetic runs in the sm_100a
kernels of libtasr_kernels.so through engine.ConformerEngine; there is no PyTorch/CPU fallback."""
import torch
import torch.nn as nn

from .attention import RelativeMultiHeadAttention


class SwiGLUFeedForward(nn.Module):
    """reference model/conformer.py:8-26."""

    def __init__(self, d_model, dim_feedforward, dropout=0.1):
        super().__init__()
        self.linear1 = nn.Linear(d_model, 2 * dim_feedforward)
        self.dropout1 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, x):
        from ..functional import ff_module_forward
        return ff_module_forward(self, x)


class TransposeGroupNorm(nn.Module):
    """reference model/conformer.py:28-49 (GroupNorm over channels-in-group x time on (N, L, C) input)."""

    def __init__(self, num_channels, num_groups=32):
        super().__init__()
        if num_channels % num_groups != 0:
            num_groups = 1
            for i in [32, 16, 8, 4, 2]:
                if num_channels % i == 0:
                    num_groups = i
                    break
        self.norm = nn.GroupNorm(num_groups=num_groups, num_channels=num_channels)

    def forward(self, x):
        from ..functional import groupnorm_module_forward
        return groupnorm_module_forward(self, x)


class ConformerConvModule(nn.Module):
    """reference model/conformer.py:51-88."""

    def __init__(self, d_model, kernel_size=31):
        super().__init__()
        if kernel_size != 31:
            raise NotImplementedError("the B200 depthwise kernel implements the reference kernel_size=31")
        self.norm = TransposeGroupNorm(d_model)
        self.pointwise_conv1 = nn.Conv1d(d_model, 2 * d_model, kernel_size=1)
        self.glu = nn.GLU(dim=1)
        self.depthwise_conv = nn.Conv1d(d_model, d_model, kernel_size=kernel_size, padding=(kernel_size - 1) // 2,
                                        groups=d_model)
        self.batch_norm = nn.BatchNorm1d(d_model)
        self.swish = nn.SiLU()
        self.pointwise_conv2 = nn.Conv1d(d_model, d_model, kernel_size=1)

    def forward(self, x):
        from ..functional import conv_module_forward
        return conv_module_forward(self, x)


class ConformerBlock(nn.Module):
    """reference model/conformer.py:90-135."""

    def __init__(self, d_model, n_heads, dropout=0.1):
        super().__init__()
        self.ff1 = SwiGLUFeedForward(d_model, d_model * 4, dropout)
        self.norm_ff1 = TransposeGroupNorm(d_model)
        self.attn = RelativeMultiHeadAttention(d_model, n_heads, dropout=dropout)
        self.norm_attn = TransposeGroupNorm(d_model)
        self.conv = ConformerConvModule(d_model)
        self.norm_conv = TransposeGroupNorm(d_model)  # constructed but never used, like the reference (:105)
        self.ff2 = SwiGLUFeedForward(d_model, d_model * 4, dropout)
        self.norm_ff2 = TransposeGroupNorm(d_model)
        self.final_norm = TransposeGroupNorm(d_model)
        self.dropout_p = dropout

    def forward(self, x, mask=None):
        from ..functional import block_module_forward
        return block_module_forward(self, x, mask)


class TurkishASRModel(nn.Module):
    """reference model/conformer.py:137-211.  forward(x (B,T,F), input_lengths (B,) | None) -> (B,T',n_classes)."""

    def __init__(self, n_mel_channels, d_model=256, n_heads=4, n_blocks=6, n_classes=31, dropout=0.1):
        super().__init__()
        self.subsample = nn.Sequential(
            nn.Conv2d(1, d_model, kernel_size=3, stride=2, padding=1),
            nn.SiLU(),
            nn.Conv2d(d_model, d_model, kernel_size=3, stride=2, padding=1),
            nn.SiLU(),
        )
        flattened_dim = d_model * (n_mel_channels // 4)
        self.input_proj = nn.Linear(flattened_dim, d_model)
        self.blocks = nn.ModuleList([ConformerBlock(d_model, n_heads, dropout=dropout) for _ in range(n_blocks)])
        self.fc = nn.Linear(d_model, n_classes)
        self.d_model = d_model
        self.n_heads = n_heads
        self.dropout_p = dropout
        self._engine = None
        self.register_load_state_dict_post_hook(TurkishASRModel._invalidate_operands)

    @staticmethod
    def _invalidate_operands(module, incompatible_keys):
        """load_state_dict rewrote the fp32 parameters: the engine's bf16 operand copies are stale."""
        eng = module._engine
        if eng is not None and eng.flat is not None:
            eng.flat.shadow_fresh = False

    def set_precision(self, precision: str):
        """"bf16" (default; bf16 GEMM operands, training + inference) or "fp32" (forward only; fp32 activations and
        fp32-accurate contractions, logits within 1e-4 of the reference's fp32 run)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.engine().precision = precision
        return self

    def engine(self):
        if self._engine is None:
            from ..engine import ConformerEngine
            object.__setattr__(self, "_engine", ConformerEngine(self))
        return self._engine

    def forward(self, x, input_lengths=None):
        from ..functional import model_forward
        return model_forward(self, x, input_lengths)
