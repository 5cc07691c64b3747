"""Drop-in mirror of the reference's quantizer library (models_fp_quant*/quant_utils.py) on the
B200 kernels.

Same function and class names, argument meaning, output dtypes and error behaviour as the
reference ("qu.py" = models_fp_quant_transform_rotate/quant_utils.py, "qu0.py" =
models_fp_quant/quant_utils.py); every FP quantizer is ONE fused kernel launch through the C ABI
(include/fpq_b200.h) instead of ~12-25 ATen launches + quant_cuda.quant.  CUDA tensors only: there
is no CPU or PyTorch fallback for the quantizers (a CPU tensor raises FpqError).

What is NOT here, and why: the integer RTN / log2 baselines
(`quantize_*_sym/asymmetric`, `log2_quant_*`, qu.py:12-206) are outside the hot path; selecting them
raises NotImplementedError naming the reference function.

Deviation from the reference, on purpose: `from_float` with a `weight_quant` that is none of
per_channel / per_tensor / per_group raises ValueError (the reference silently keeps the random
fp16 weight the constructor allocated, qu.py:677-685,794-855).
"""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

from . import ops
from ._lib import FpqError

__all__ = [
    "quantize_to_nearest_grid", "fp4_e3m0_grid", "fp4_e2m1_grid", "fp4_e1m2_grid", "fp6_e2m3_grid", "fp6_e3m2_grid",
    "int_neg_grid", "e2m3_pos_grid",
    "fp_quant_e1_per_token", "fp_quant_e2_per_token", "fp_quant_e3_per_token",
    "fp_quant_e1_per_group", "fp_quant_e2_per_group", "fp_quant_e3_per_group",
    "fp_quant_e1_per_group_cuda", "fp_quant_e2_per_group_cuda", "fp_quant_e3_per_group_cuda",
    "fp_quant_e1m2_neg_e2m1_pos_per_group", "fp_quant_e1m2_neg_e2m1_pos_per_group_cuda",
    "fp6_quant_e2m3_per_token_cuda", "fp6_quant_e3m2_per_token_cuda", "fp6_quant_e2m3_per_group_cuda", "fp6_quant_e3m2_per_group_cuda",
    "fp6_quant_int_neg_e2m3_pos_per_group_cuda", "fp6_quant_int_neg_e2m3_pos_per_token_cuda",
    "fp4_afpq_per_group_cuda", "fp_neg_reverse_quant_per_group_cuda",
    "QuantizedLinear", "QuantizedLinear_fc2", "quantize_VAR",
]

# ----------------------------------------------------------------------------------------------
# grids (qu.py:233-235, 458-500) -- CPU tensors, like the reference's module-level constants
# ----------------------------------------------------------------------------------------------
def _sym(pos, double_zero=False):
    return torch.tensor([-v for v in reversed(pos)] + ([0.0, 0.0] if double_zero else [0.0]) + list(pos), dtype=torch.float32)


_E2M3_POS = [0.125 * i for i in range(1, 16)] + [2.0 + 0.25 * i for i in range(8)] + [4.0 + 0.5 * i for i in range(8)]
_E3M2_POS = ([0.0625 * i for i in range(1, 8)] + [0.5 + 0.125 * i for i in range(4)] + [1.0 + 0.25 * i for i in range(4)]
             + [2.0 + 0.5 * i for i in range(4)] + [4.0 + 1.0 * i for i in range(4)] + [8.0 + 2.0 * i for i in range(4)]
             + [16.0 + 4.0 * i for i in range(4)])
fp4_e3m0_grid = _sym([0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0])
fp4_e2m1_grid = _sym([0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
fp4_e1m2_grid = _sym([0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75])
fp6_e2m3_grid = _sym(_E2M3_POS, double_zero=True)
fp6_e3m2_grid = _sym(_E3M2_POS, double_zero=True)
int_neg_grid = torch.tensor([float(-i) for i in range(32, -1, -1)], dtype=torch.float32)
e2m3_pos_grid = torch.tensor([0.0] + _E2M3_POS, dtype=torch.float32)


def quantize_to_nearest_grid(x: torch.Tensor, quant_grid: torch.Tensor):
    """qu.py:209-230: nearest grid value by argmin over |x - grid| (first minimal index: exact ties go to
    the smaller value, NaN -> grid[0]); result in the grid's dtype."""
    grid = quant_grid.to(device=x.device, dtype=torch.float32)
    out = ops.quant_grid(x.to(torch.float32).reshape(-1), grid, "argmin").view(x.shape)
    return out.to(quant_grid.dtype)


# ----------------------------------------------------------------------------------------------
# FP4 symmetric
# ----------------------------------------------------------------------------------------------
def _per_token(x, n_bits, fmt):
    assert n_bits == 4
    return ops.fake_quant(x, fmt, None, "argmin", clamp3=True)          # qu.py:237-247: clamp +-3, scale over the last dim, fp32 result


def fp_quant_e3_per_token(x, n_bits):
    return _per_token(x, n_bits, "e3m0")


def fp_quant_e2_per_token(x, n_bits):
    return _per_token(x, n_bits, "e2m1")


def fp_quant_e1_per_token(x, n_bits):
    return _per_token(x, n_bits, "e1m2")


def fp_quant_e3_per_group(x, n_bits, group_size=128):
    assert n_bits == 4
    return ops.fake_quant(x, "e3m0", group_size, "argmin", clamp3=True)   # qu.py:250-262


def fp_quant_e1_per_group(x, n_bits, group_size=128):
    assert n_bits == 4
    return ops.fake_quant(x, "e1m2", group_size, "argmin", clamp3=True)   # qu.py:346-358


def fp_quant_e2_per_group(x, n_bits, group_size=128):
    """qu.py:298-310.  No clamp, and -- as in the reference -- the caller's tensor is left holding
    the NORMALISED values (`x.view(-1, g).div_(scale)`, qu.py:306): the side effect is reproduced."""
    assert n_bits == 4
    out = ops.fake_quant(x, "e2m1", group_size, "argmin")
    xv = x.view(-1, group_size)                                           # raises for non-contiguous input, as the reference's view does
    xv.div_(xv.abs().max(dim=-1, keepdim=True)[0] / fp4_e2m1_grid.abs().max())
    return out


def _group_cuda(x, n_bits, want_bits, fmt, group_size, out_dtype=None):
    assert n_bits == want_bits
    return ops.fake_quant(x, fmt, group_size, "kernel", out_dtype=out_dtype)


def fp_quant_e3_per_group_cuda(x, n_bits, group_size=128):
    return _group_cuda(x, n_bits, 4, "e3m0", group_size)                  # qu.py:265-282


def fp_quant_e2_per_group_cuda(x, n_bits, group_size=128):
    return _group_cuda(x, n_bits, 4, "e2m1", group_size)                  # qu.py:313-330


def fp_quant_e1_per_group_cuda(x, n_bits, group_size=128):
    return _group_cuda(x, n_bits, 4, "e1m2", group_size)                  # qu.py:361-378


# ----------------------------------------------------------------------------------------------
# FP4 sign-split (fc2 inputs)
# ----------------------------------------------------------------------------------------------
def _global_clip(x, clipping_strength):
    """qu.py:421-422 for a strength other than 1.0 (at 1.0 the kernel reproduces it without a pass)."""
    clip_value = clipping_strength * x.abs().max()
    return torch.clamp(x, -clip_value, clip_value)


def fp_quant_e1m2_neg_e2m1_pos_per_group(x, n_bits, group_size=128, clipping_strength=1.0):
    assert n_bits == 4                                                     # qu.py:381-412
    if clipping_strength != 1.0:
        return ops.fake_quant_signsplit(_global_clip(x, clipping_strength), "e1m2_neg_e2m1_pos", group_size, "argmin", global_clip=True)
    return ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", group_size, "argmin", global_clip=True)


def fp_quant_e1m2_neg_e2m1_pos_per_group_cuda(x, n_bits, group_size=128, clipping_strength=1.0):
    assert n_bits == 4                                                     # qu.py:415-452
    if clipping_strength != 1.0:
        x = _global_clip(x, clipping_strength)
    return ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", group_size, "kernel", global_clip=True)


def fp4_afpq_per_group_cuda(x, n_bits, group_size=128, clipping_strength=1.0):
    assert n_bits == 4                                                     # qu0.py:498-535
    if clipping_strength != 1.0:
        x = _global_clip(x, clipping_strength)
    return ops.fake_quant_signsplit(x, "afpq_e2m1", group_size, "kernel", global_clip=True)


def fp_neg_reverse_quant_per_group_cuda(x, n_bits, group_size=128):
    """qu0.py:454-495 (an ablation format of models_fp_quant): negatives are shifted by |min| of their
    group and quantized on the full e2m1 grid, positives on e2m1; the shift is taken off every element.
    Composed from the element-rule kernel (fpq_quant_grid) and ATen glue, like the reference composes
    it around quant_cuda.quant; not a fused kernel (it is on no README path)."""
    assert n_bits == 4
    grid = fp4_e2m1_grid.to(x.device)
    shape = x.shape
    g = x.reshape(-1, group_size)
    m = g.min(dim=-1, keepdim=True)[0].abs()
    zero = torch.zeros_like(g)
    xr = torch.where(g <= 0, g, zero) + m
    xp = torch.where(g > 0, g, zero)
    sr = xr.abs().max(dim=-1, keepdim=True)[0] / grid.abs().max()
    sp = xp.abs().max(dim=-1, keepdim=True)[0] / grid.abs().max()
    qr = ops.quant_grid((xr / sr).view(-1).to(torch.float32), grid, "kernel").view(g.shape)
    qp = ops.quant_grid((xp / sp).view(-1).to(torch.float32), grid, "kernel").view(g.shape)
    out = (qr * sr - m) + qp * sp
    return out.view(shape).to(x.dtype)


# ----------------------------------------------------------------------------------------------
# FP6
# ----------------------------------------------------------------------------------------------
def fp6_quant_e2m3_per_token_cuda(x, n_bits):
    assert n_bits == 6
    return ops.fake_quant(x, "e2m3", None, "kernel", out_dtype=torch.float16)      # qu.py:503-517: always fp16


def fp6_quant_e3m2_per_token_cuda(x, n_bits):
    assert n_bits == 6
    return ops.fake_quant(x, "e3m2", None, "kernel", out_dtype=torch.float16)      # qu.py:520-534


def fp6_quant_e2m3_per_group_cuda(x, n_bits, group_size=128):
    return _group_cuda(x, n_bits, 6, "e2m3", group_size, torch.float16)            # qu.py:537-554


def fp6_quant_e3m2_per_group_cuda(x, n_bits, group_size=128):
    return _group_cuda(x, n_bits, 6, "e3m2", group_size, torch.float16)            # qu.py:557-574


def fp6_quant_int_neg_e2m3_pos_per_group_cuda(x, n_bits, group_size=128):
    assert n_bits == 6
    return ops.fake_quant_signsplit(x, "int_neg_e2m3_pos", group_size, "kernel")   # qu.py:577-611 (no global clip)


def fp6_quant_int_neg_e2m3_pos_per_token_cuda(x, n_bits):
    assert n_bits == 6
    return ops.fake_quant_signsplit(x, "int_neg_e2m3_pos", None, "kernel")         # qu.py:614-646


# ----------------------------------------------------------------------------------------------
# out-of-scope baselines
# ----------------------------------------------------------------------------------------------
def _out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(f"{name} (integer RTN / log2 baseline of the reference's quant_utils.py) is outside the FP "
                                  "fake-quant hot path this package implements; use the reference for that baseline")
    fn.__name__ = name
    return fn


quantize_activation_per_token_sym = _out_of_scope("quantize_activation_per_token_sym")
quantize_activation_per_token_asymmetric = _out_of_scope("quantize_activation_per_token_asymmetric")
quantize_activation_per_tensor_sym = _out_of_scope("quantize_activation_per_tensor_sym")
quantize_activation_per_tensor_asymmetric = _out_of_scope("quantize_activation_per_tensor_asymmetric")
quantize_activation_per_group_sym = _out_of_scope("quantize_activation_per_group_sym")
quantize_activation_per_group_asymmetric = _out_of_scope("quantize_activation_per_group_asymmetric")
log2_quant_per_group_asym = _out_of_scope("log2_quant_per_group_asym")
log2_quant_per_token_asym = _out_of_scope("log2_quant_per_token_asym")
quantize_weight_per_channel_sym = _out_of_scope("quantize_weight_per_channel_sym")
quantize_weight_per_tensor_sym = _out_of_scope("quantize_weight_per_tensor_sym")
quantize_weight_per_group_sym = _out_of_scope("quantize_weight_per_group_sym")


# ----------------------------------------------------------------------------------------------
# QuantizedLinear / QuantizedLinear_fc2 / quantize_VAR
# ----------------------------------------------------------------------------------------------
_PER_TOKEN_FP = {                                        # qu.py:695-709
    "fp_e1": fp_quant_e1_per_token, "fp_e2": fp_quant_e2_per_token, "fp_e3": fp_quant_e3_per_token,
    "fp6_e2m3": fp6_quant_e2m3_per_token_cuda, "fp6_e3m2": fp6_quant_e3m2_per_token_cuda,
}
_PER_GROUP_FP = {                                        # qu.py:724-737
    "fp_e1": fp_quant_e1_per_group_cuda, "fp_e2": fp_quant_e2_per_group_cuda, "fp_e3": fp_quant_e3_per_group_cuda,
    "fp6_e2m3": fp6_quant_e2m3_per_group_cuda, "fp6_e3m2": fp6_quant_e3m2_per_group_cuda,
}
_FC2_PER_TOKEN_EXTRA = {"fp6_int_neg_e2m3_pos": fp6_quant_int_neg_e2m3_pos_per_token_cuda}
_FC2_PER_GROUP_EXTRA = {                                 # qu.py:955-962, qu0.py:1038-1041
    "fp_e1m2_neg_e2m1_pos": fp_quant_e1m2_neg_e2m1_pos_per_group_cuda,
    "fp6_int_neg_e2m3_pos": fp6_quant_int_neg_e2m3_pos_per_group_cuda,
    "fp_neg_reverse_quant": fp_neg_reverse_quant_per_group_cuda,
    "fp4_afpq": fp4_afpq_per_group_cuda,
}


def _identity(x):
    return x


class _QuantizedLinearBase(nn.Module):
    _NAME = "QuantizedLinear"
    _PER_TOKEN = _PER_TOKEN_FP
    _PER_GROUP = _PER_GROUP_FP

    def __init__(self, in_features, out_features, bias=True, act_quant=None, quantize_output=False, w_bit=8, a_bit=8,
                 act_quant_sym=True, fc2_act_log2_quant=False, activation_fp_quant=False, weight_fp_quant=False,
                 act_fp_type=False, weight_fp_type=False):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.w_bit = w_bit
        self.a_bit = a_bit
        self.act_quant_sym = act_quant_sym
        self.fc2_act_log2_quant = fc2_act_log2_quant
        self.activation_fp_quant = activation_fp_quant
        self.weight_fp_quant = weight_fp_quant
        self.act_fp_type = act_fp_type
        self.weight_quant_name = None
        self.register_buffer("weight", torch.randn(self.out_features, self.in_features, dtype=torch.float16, requires_grad=False))
        if bias:
            self.register_buffer("bias", torch.zeros((1, self.out_features), dtype=torch.float16, requires_grad=False))
        else:
            self.register_buffer("bias", None)

        if act_quant == "per_token":
            self.act_quant_name = "per_token"
            if self.activation_fp_quant == True:  # noqa: E712  (the reference compares with ==)
                self.act_quant = partial(self._lookup(self._PER_TOKEN, act_fp_type), n_bits=a_bit)
            elif self.act_quant_sym == True:  # noqa: E712
                self.act_quant = partial(quantize_activation_per_token_sym, n_bits=a_bit)
            else:
                self.act_quant = partial(quantize_activation_per_token_asymmetric, n_bits=a_bit)
        elif act_quant == "per_tensor":
            self.act_quant_name = "per_tensor"
            if self.act_quant_sym == True:  # noqa: E712
                self.act_quant = partial(quantize_activation_per_tensor_sym, n_bits=a_bit)
            else:
                self.act_quant = partial(quantize_activation_per_tensor_asymmetric, n_bits=a_bit)
        elif act_quant == "per_group":
            self.act_quant_name = "per_group"
            if self.activation_fp_quant == True:  # noqa: E712
                self.act_quant = partial(self._lookup(self._PER_GROUP, act_fp_type), n_bits=a_bit, group_size=128)
            elif self.fc2_act_log2_quant == True:  # noqa: E712
                self.act_quant = partial(log2_quant_per_group_asym, n_bits=a_bit, group_size=128)
            elif self.act_quant_sym == True:  # noqa: E712
                self.act_quant = partial(quantize_activation_per_group_sym, n_bits=a_bit, group_size=128)
            else:
                self.act_quant = partial(quantize_activation_per_group_asymmetric, n_bits=a_bit, group_size=128)
        else:
            raise ValueError(f"Invalid act_quant: {act_quant}")

        if quantize_output:
            self.output_quant_name = self.act_quant_name
            self.output_quant = self.act_quant
        else:
            self.output_quant_name = "None"
            self.output_quant = _identity

    @staticmethod
    def _lookup(table, fp_type):
        try:
            return table[fp_type]
        except (KeyError, TypeError):
            raise ValueError("Unsupported fp_type.") from None            # qu.py:709,737

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.weight = self.weight.to(*args, **kwargs)
        if self.bias is not None:
            self.bias = self.bias.to(*args, **kwargs)
        return self

    @torch.no_grad()
    def forward(self, x):                                                 # qu.py:764-769 / :991-996
        q_x = self.act_quant(x)
        y = torch.functional.F.linear(q_x, self.weight, self.bias)
        return self.output_quant(y)

    @classmethod
    def _from_float(cls, module, weight_quant, act_quant, quantize_output, w_bit, a_bit, act_quant_sym, fc2_act_log2_quant,
                    activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type):
        assert isinstance(module, torch.nn.Linear)
        new_module = cls(module.in_features, module.out_features, module.bias is not None, act_quant=act_quant,
                         quantize_output=quantize_output, w_bit=w_bit, a_bit=a_bit, act_quant_sym=act_quant_sym,
                         fc2_act_log2_quant=fc2_act_log2_quant, activation_fp_quant=activation_fp_quant,
                         weight_fp_quant=weight_fp_quant, act_fp_type=act_fp_type, weight_fp_type=weight_fp_type)
        w = module.weight.detach()
        if weight_quant == "per_channel":                                 # qu.py:794-818
            if weight_fp_quant == True:  # noqa: E712
                new_module.weight = cls._lookup(_PER_TOKEN_FP, weight_fp_type)(w, n_bits=w_bit)
            else:
                new_module.weight = quantize_weight_per_channel_sym(w, n_bits=w_bit)
        elif weight_quant == "per_tensor":                                # qu.py:820-823
            new_module.weight = quantize_weight_per_tensor_sym(w, n_bits=w_bit)
        elif weight_quant == "per_group":                                 # qu.py:825-855
            if weight_fp_quant == True:  # noqa: E712
                new_module.weight = cls._lookup(_PER_GROUP_FP, weight_fp_type)(w, n_bits=w_bit, group_size=128)
            else:
                new_module.weight = quantize_weight_per_group_sym(w, n_bits=w_bit, group_size=128)
        else:
            raise ValueError(f"Invalid weight_quant: {weight_quant}")
        new_module.weight_quant_name = weight_quant
        if module.bias is not None:
            new_module.bias = module.bias                                 # qu.py:858-859: the original Parameter
        return new_module


class QuantizedLinear(_QuantizedLinearBase):
    """qu.py:649-867."""

    @staticmethod
    def from_float(module, weight_quant="per_channel", act_quant="per_token", quantize_output=False, w_bit=8, a_bit=8,
                   act_quant_sym=None, fc2_act_log2_quant=False, activation_fp_quant=False, weight_fp_quant=False,
                   act_fp_type=None, weight_fp_type=None):
        return QuantizedLinear._from_float(module, weight_quant, act_quant, quantize_output, w_bit, a_bit, act_quant_sym,
                                           fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type)

    def __repr__(self):
        return f"QuantizedLinear{self.in_features}, {self.out_features}, bias={self.bias is not None}, " \
            f"weight_quant={self.weight_quant_name}, act_quant={self.act_quant_name}, output_quant={self.output_quant_name}, " \
            f"w_bit={self.w_bit}, a_bit={self.a_bit}, act_quant_sym={self.act_quant_sym}, act_log2_quant={self.fc2_act_log2_quant}," \
            f"activation_fp_quant={self.activation_fp_quant}, weight_fp_quant={self.weight_fp_quant}, " \
            f"activation_quant_type={self.act_fp_type}"


class QuantizedLinear_fc2(_QuantizedLinearBase):
    """qu.py:870-1093: the fc2 variant accepts the sign-split activation formats."""
    _NAME = "QuantizedLinear_fc2"
    _PER_TOKEN = {**_PER_TOKEN_FP, **_FC2_PER_TOKEN_EXTRA}
    _PER_GROUP = {**_PER_GROUP_FP, **_FC2_PER_GROUP_EXTRA}

    @staticmethod
    def from_float(module, weight_quant="per_channel", act_quant="per_token", quantize_output=False, w_bit=8, a_bit=8,
                   act_quant_sym=None, fc2_act_log2_quant=False, activation_fp_quant=False, weight_fp_quant=False,
                   act_fp_type=None, weight_fp_type=None):
        return QuantizedLinear_fc2._from_float(module, weight_quant, act_quant, quantize_output, w_bit, a_bit, act_quant_sym,
                                               fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type)

    def __repr__(self):
        return f"QuantizedLinear_fc2{self.in_features}, {self.out_features}, bias={self.bias is not None}, " \
            f"weight_quant={self.weight_quant_name}, act_quant={self.act_quant_name}, output_quant={self.output_quant_name}, " \
            f"w_bit={self.w_bit}, a_bit={self.a_bit}, act_quant_sym={self.act_quant_sym}, act_log2_quant={self.fc2_act_log2_quant}," \
            f"activation_fp_quant={self.activation_fp_quant}, weight_fp_quant={self.weight_fp_quant}"


def _is_ffn(m):
    return type(m).__name__ == "FFN" and isinstance(getattr(m, "fc1", None), nn.Linear) and isinstance(getattr(m, "fc2", None), nn.Linear)


def _is_self_attention(m):
    return type(m).__name__ == "SelfAttention" and isinstance(getattr(m, "mat_qkv", None), nn.Linear) \
        and isinstance(getattr(m, "proj", None), nn.Linear)


def quantize_VAR(model, weight_quant=None, act_quant=None, quantize_bmm_input=False, w_bit=8, a_bit=8, kv_bit=8,
                 act_quant_sym=None, fc2_act_log2_quant=None, quant_kv=None, activation_fp_quant=False, weight_fp_quant=False,
                 act_fp_type=None, weight_fp_type=None, fc2_fp_type=None):
    """qu.py:1095-1166: swap fc1 / mat_qkv / proj -> QuantizedLinear and fc2 -> QuantizedLinear_fc2 in every
    FFN / SelfAttention module.  Modules are recognised by class NAME and attributes, so the function
    works on the reference's model classes (any of its five model packages) without importing them."""
    common = dict(weight_quant=weight_quant, act_quant=act_quant, w_bit=w_bit, a_bit=a_bit,
                  activation_fp_quant=activation_fp_quant, weight_fp_quant=weight_fp_quant, weight_fp_type=weight_fp_type)
    for _name, m in list(model.named_modules()):
        if _is_ffn(m):
            m.fc1 = QuantizedLinear.from_float(m.fc1, act_quant_sym=act_quant_sym, act_fp_type=act_fp_type, **common)
            m.fc2 = QuantizedLinear_fc2.from_float(m.fc2, act_quant_sym=False, fc2_act_log2_quant=fc2_act_log2_quant,
                                                   act_fp_type=fc2_fp_type, **common)
        elif _is_self_attention(m):
            m.mat_qkv = QuantizedLinear.from_float(m.mat_qkv, act_quant_sym=act_quant_sym, act_fp_type=act_fp_type, **common)
            m.proj = QuantizedLinear.from_float(m.proj, act_quant_sym=act_quant_sym, act_fp_type=act_fp_type, **common)
    return model


# ---- per-layer ("mixed datatype") variants ---------------------------------------------------------------------
# evaluate_fp_quant.py:18 imports quantize_VAR_mixed_fp4_datatype / quantize_VAR_mixed_fp6_datatype next to quantize_VAR
# (models_fp_quant/quant_utils.py:1256-1432); models_fp_quant_rotate/quant_utils.py:982-1067 has
# quantize_VAR_use_different_datatype.  They differ from quantize_VAR in two ways: the (activation, weight) format of a
# linear depends on its block index (tables taken from the format search), and the adaLN projection `ada_lin[1]` of every
# block is quantized as well.  A plan maps (site, block) to formats; `None` entries fall back to the caller's arguments.
_FC1_E2_BLOCKS = frozenset(range(6, 21))                     # fc1 activations on fp_e2 in blocks 6..20, fp_e3 elsewhere


def _plan_mixed_fp4(qkv_e2_blocks):
    def plan(site, block):
        if site == "fc1":
            return ("fp_e2" if block in _FC1_E2_BLOCKS else "fp_e3", "fp_e2")
        if site == "mat_qkv":
            return ("fp_e2" if block in qkv_e2_blocks else "fp_e3", "fp_e2")
        return (None, None)                                     # proj, fc2, ada_lin: the caller's act_fp_type / fc2_fp_type / weight_fp_type
    return plan


def _plan_mixed_fp6(site, block):
    if site in ("fc1", "mat_qkv"):
        return ("fp6_e3m2", "fp6_e2m3")
    if site == "fc2":
        return ("fp6_e2m3" if block in (0, 23) else "fp6_e3m2", "fp6_e2m3")
    if site == "proj":
        return ("fp6_e2m3" if block >= 2 else "fp6_e3m2", "fp6_e2m3")
    return ("fp6_e2m3", "fp6_e2m3")                             # ada_lin


def _quantize_VAR_planned(model, plan, weight_quant, act_quant, w_bit, a_bit, act_quant_sym, fc2_act_log2_quant,
                          activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type, fc2_fp_type):
    common = dict(weight_quant=weight_quant, act_quant=act_quant, w_bit=w_bit, a_bit=a_bit,
                  activation_fp_quant=activation_fp_quant, weight_fp_quant=weight_fp_quant)

    def swap(owner, attr, site, block, fc2=False):
        a, w = plan(site, block)
        a = a if a is not None else (fc2_fp_type if fc2 else act_fp_type)
        w = w if w is not None else weight_fp_type
        target = owner[attr] if isinstance(attr, int) else getattr(owner, attr)
        if fc2:
            new = QuantizedLinear_fc2.from_float(target, act_quant_sym=False, fc2_act_log2_quant=fc2_act_log2_quant,
                                                 act_fp_type=a, weight_fp_type=w, **common)
        else:
            new = QuantizedLinear.from_float(target, act_quant_sym=act_quant_sym, act_fp_type=a, weight_fp_type=w, **common)
        if isinstance(attr, int):
            owner[attr] = new
        else:
            setattr(owner, attr, new)

    for name, m in list(model.named_modules()):             # parents come before their children, as in the reference's walk
        if not (_is_ffn(m) or _is_self_attention(m) or type(m).__name__ == "AdaLNSelfAttn"):
            continue
        block = int(name.split(".")[1])                        # "blocks.<i>...." (qu0.py:1270)
        if _is_ffn(m):
            swap(m, "fc1", "fc1", block)
            swap(m, "fc2", "fc2", block, fc2=True)
        elif _is_self_attention(m):
            swap(m, "mat_qkv", "mat_qkv", block)
            swap(m, "proj", "proj", block)
        else:
            swap(m.ada_lin, 1, "ada_lin", block)               # AttributeError with shared_aln=True, as in the reference
    return model


def quantize_VAR_with_ada_lin(model, weight_quant=None, act_quant=None, quantize_bmm_input=False, w_bit=8, a_bit=8, kv_bit=8,
                              act_quant_sym=None, fc2_act_log2_quant=None, quant_kv=None, activation_fp_quant=False,
                              weight_fp_quant=False, act_fp_type=None, weight_fp_type=None, fc2_fp_type=None):
    """`quantize_VAR` as the `models_fp_quant_rotate` package defines it (models_fp_quant_rotate/quant_utils.py:894-979):
    the four linears of every block as in `quantize_VAR`, plus the adaLN projection `ada_lin[1]` of every block, which
    the other packages leave in floating point (their branch is commented out, qu.py:1147-1155)."""
    return _quantize_VAR_planned(model, lambda site, block: (None, None), weight_quant, act_quant, w_bit, a_bit, act_quant_sym,
                                 fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type, fc2_fp_type)


def quantize_VAR_mixed_fp4_datatype(model, weight_quant=None, act_quant=None, quantize_bmm_input=False, w_bit=8, a_bit=8, kv_bit=8,
                                    act_quant_sym=None, fc2_act_log2_quant=None, quant_kv=None, activation_fp_quant=False,
                                    weight_fp_quant=False, act_fp_type=None, weight_fp_type=None, fc2_fp_type=None):
    """models_fp_quant/quant_utils.py:1256-1341."""
    return _quantize_VAR_planned(model, _plan_mixed_fp4(frozenset((0, 24, 25))), weight_quant, act_quant, w_bit, a_bit, act_quant_sym,
                                 fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type, fc2_fp_type)


def quantize_VAR_mixed_fp6_datatype(model, weight_quant=None, act_quant=None, quantize_bmm_input=False, w_bit=8, a_bit=8, kv_bit=8,
                                    act_quant_sym=None, fc2_act_log2_quant=None, quant_kv=None, activation_fp_quant=False,
                                    weight_fp_quant=False, act_fp_type=None, weight_fp_type=None, fc2_fp_type=None):
    """models_fp_quant/quant_utils.py:1344-1432 (every format comes from the table; the *_fp_type arguments are unused there too)."""
    return _quantize_VAR_planned(model, _plan_mixed_fp6, weight_quant, act_quant, w_bit, a_bit, act_quant_sym,
                                 fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type, fc2_fp_type)


def quantize_VAR_use_different_datatype(model, weight_quant=None, act_quant=None, quantize_bmm_input=False, w_bit=8, a_bit=8, kv_bit=8,
                                        act_quant_sym=None, fc2_act_log2_quant=None, quant_kv=None, activation_fp_quant=False,
                                        weight_fp_quant=False, act_fp_type=None, weight_fp_type=None, fc2_fp_type=None):
    """models_fp_quant_rotate/quant_utils.py:982-1067 (the FP4 table without block 0 in the mat_qkv list)."""
    return _quantize_VAR_planned(model, _plan_mixed_fp4(frozenset((24, 25))), weight_quant, act_quant, w_bit, a_bit, act_quant_sym,
                                 fc2_act_log2_quant, activation_fp_quant, weight_fp_quant, act_fp_type, weight_fp_type, fc2_fp_type)


assert FpqError  # re-exported for callers that want to catch it
