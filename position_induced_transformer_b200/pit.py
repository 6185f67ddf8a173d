"""Drop-in replacement for the reference module ``pit.py`` with fused sm_100a position-attention.

Same public names, constructor signatures, attribute names, ``state_dict`` keys and
random-number consumption at construction as the reference, so the ``train_*.py`` scripts
(``from pit import *``) run unchanged.  What differs is inside the ``posatt*`` layers:
``forward`` calls one fused CUDA op (``posatt.position_attention``) instead of materialising
the distance matrix, sorting every row for the quantile, masking, soft-maxing and contracting
(pit.py:37-57 and the three variant families at pit.py:129-159, 186-215, 243-273).

Like the reference (pit.py:1-11) this module re-exports ``torch, nn, gelu, np, pi`` and applies
the same process-wide settings at import.
"""
import torch

torch.set_float32_matmul_precision("high")  # pit.py:2 -- governs the kaiming_mlp Linears (cuBLAS TF32)
torch.manual_seed(0)
torch.cuda.manual_seed(0)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.deterministic = True
import numpy as np
import torch.nn as nn
from torch.nn.functional import gelu

np.random.seed(0)
from math import pi

from .posatt import (bias_act, bias_act_supported, decoder_tail, decoder_tail_supported, head_scale_cuda, meshes_need_grad,
                     mlp_fused, mlp_fused_supported, position_attention, processor_blocks, processor_supported)

__all__ = [
    "torch", "nn", "gelu", "np", "pi", "kaiming_mlp", "use_host_scale_map", "use_fused_decoder_tail", "use_fused_mlp_epilogue",
    "use_fused_processor",
    "posatt", "posatt_cross", "pit",
    "posatt_fixed", "posatt_cross_fixed", "pit_fixed",
    "posatt_periodic1d", "posatt_cross_periodic1d", "pit_periodic1d",
    "posatt_periodic2d", "posatt_cross_periodic2d", "pit_periodic2d",
]

# 0.25*pi*(1-1e-7) of pit.py:48: a python double that torch rounds to fp32 inside the op.
_SCALE_CONST = 0.25 * pi * (1 - 1e-7)


_HOST_SCALE_MAP = False
_FUSED_MLP_EPILOGUE = True


def use_fused_mlp_epilogue(enabled: bool) -> None:
    """Switch the fused bias/GELU epilogues of kaiming_mlp on or off; off runs the Linears and GELUs as separate torch ops."""
    global _FUSED_MLP_EPILOGUE
    _FUSED_MLP_EPILOGUE = bool(enabled)



def use_host_scale_map(enabled: bool) -> None:
    """Parity-testing aid.  The locality mask is decided on values that tie to within one ulp, so the
    last bit of s_h matters: CUDA's sin/tan and the CPU's differ in that bit for
    some lmda, which flips mask entries exactly as it does between the reference's own GPU and CPU runs.
    With the host map enabled the H scalars are evaluated with the CPU's libm (one device->host sync per
    stage) so that outputs can be compared with a CPU run of the reference to 1e-5."""
    global _HOST_SCALE_MAP
    _HOST_SCALE_MAP = bool(enabled)


def head_scale(lmda: torch.Tensor) -> torch.Tensor:
    """Per-head positive scale s_h = tan(c * (1 + sin(lmda_h))) (pit.py:48); differentiable torch glue."""
    if _HOST_SCALE_MAP and lmda.is_cuda:
        return head_scale(lmda.cpu()).to(lmda.device)
    if lmda.is_cuda and lmda.dtype == torch.float32:
        return head_scale_cuda(lmda)  # same arithmetic in one launch (and one for the derivative) instead of 4 + 7
    return torch.tan(_SCALE_CONST * (1.0 + torch.sin(lmda)))


_TALL_ROWS = 8192           # rows above which the weight gradient of a Linear is computed as a split-K batched GEMM
_TALL_SPLITS = 256          # at most; one slice per 512 rows


class _TallLinear(torch.autograd.Function):
    """x @ W^T for a very tall x (the decoder MLP of the per-sample-mesh models runs over B * N = 225 k points, train_naca.py:60).

    Forward and the input gradient are ordinary cuBLAS GEMMs.  The WEIGHT gradient dW = dZ^T x is a [out x rows] x [rows x in]
    product with a tiny output and a huge reduction; cuBLAS' heuristic runs it on a handful of CTAs (a 64x64-tile kernel, four
    CTAs for a 128 x 128 output: 150 us at the NACA decoder).  Here the rows are cut into 256 slices, each slice is one batch of
    a batched GEMM (256 CTAs' worth of work), and the partial gradients are summed.  The latent-grid MLPs of the same models
    (14 560 rows) get one slice per 512 rows: their weight gradients ran 14 us each on the same four-CTA kernel."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return torch.nn.functional.linear(x, weight)

    @staticmethod
    def backward(ctx, dz):
        x, weight = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = dz.matmul(weight)
        if ctx.needs_input_grad[1]:
            x2, dz2 = x.reshape(-1, x.shape[-1]), dz.reshape(-1, dz.shape[-1])
            rows = x2.shape[0]
            splits = max(2, min(_TALL_SPLITS, rows // 512))
            per = rows // splits
            main = per * splits
            dw = torch.bmm(dz2[:main].view(splits, per, -1).transpose(1, 2), x2[:main].view(splits, per, -1)).sum(0)
            if main < rows:
                dw = dw + dz2[main:].t().mm(x2[main:])
        return dx, dw


def _linear_no_bias(x, weight):
    rows = x.numel() // max(1, x.shape[-1])
    if rows >= _TALL_ROWS and x.is_cuda and torch.is_grad_enabled() and weight.requires_grad:
        return _TallLinear.apply(x, weight)
    return torch.nn.functional.linear(x, weight)


class kaiming_mlp(nn.Module):
    """Linear -> GELU -> Linear with Kaiming-normal weights (pit.py:13-26)."""

    def __init__(self, n_filters0, n_filters1, n_filters2):
        super().__init__()
        widths = (n_filters0, n_filters1, n_filters2)
        for idx in (1, 2):  # creation order fixes both the state_dict order and the RNG stream
            setattr(self, f"mlp{idx}", nn.Linear(widths[idx - 1], widths[idx]))
        for idx in (1, 2):
            nn.init.kaiming_normal_(getattr(self, f"mlp{idx}").weight)

    def forward(self, x):
        return self._run(x, False)

    def forward_gelu(self, x):
        """gelu(self(x)): the activation the callers apply to the block output (pit.py:111, 121), fused into the last epilogue."""
        return self._run(x, True)

    def _run(self, x, act_out):
        l1, l2 = self.mlp1, self.mlp2
        # CUDA path: bias-free GEMMs (cuBLAS, TF32 as pit.py:2 sets) + one epilogue kernel each -- bias, GELU and, in backward,
        # GELU' together with the bias gradient, which autograd would run as a separate 11 us column reduction
        if (_FUSED_MLP_EPILOGUE and type(l1) is nn.Linear and type(l2) is nn.Linear and l1.bias is not None and l2.bias is not None
                and x.is_cuda and x.dtype == torch.float32 and x.numel() > 0):
            if mlp_fused_supported(x, l1.weight, l2.weight):       # narrow input (the encoder lift): the whole MLP in one launch
                return mlp_fused(x, l1.weight, l1.bias, l2.weight, l2.bias, act_out)
            z1 = _linear_no_bias(x, l1.weight)
            if bias_act_supported(z1, l1.bias):
                h = bias_act(z1, l1.bias, True)
                z2 = _linear_no_bias(h, l2.weight)
                if bias_act_supported(z2, l2.bias):
                    return bias_act(z2, l2.bias, act_out)
                y = z2 + l2.bias
                return gelu(y) if act_out else y
        y = l2(gelu(l1(x)))
        return gelu(y) if act_out else y


class posatt(nn.Module):
    """Position-attention over per-sample meshes; self stage (pit.py:28-57).

    Subclasses only change three class-level facts: the distance ``_variant``, whether meshes
    are shared by the batch, and whether ``forward`` is the cross form.
    """

    _variant = "euclid"
    _cross = False

    def __init__(self, n_head, in_dim, locality):
        super().__init__()
        self.locality = locality
        self.n_head = n_head
        self.in_dim = in_dim
        self.lmda = nn.Parameter(torch.rand(n_head, 1, 1))

    # -- fused path ------------------------------------------------------------------
    def _attend(self, mesh_out, mesh_in, inputs, self_concat):
        return position_attention(mesh_out, mesh_in, inputs, head_scale(self.lmda), self.locality,
                                  variant=self._variant, self_concat=self_concat)

    def forward(self, *args):
        if self._cross:
            mesh_out, mesh_in, inputs = args
            return self._attend(mesh_out, mesh_in, inputs, False)
        mesh, inputs = args
        return self._attend(mesh, mesh, inputs, True)

    # -- inspection helpers kept for API parity (never used by forward) ----------------
    def dist2att(self, mesh_out, mesh_in, scale, locality):
        """Dense attention weights ([B,]H,N,M), produced by the fused kernel applied to an identity
        value matrix -- for inspection only; ``forward`` never materialises this tensor."""
        m = mesh_in.shape[-2]
        eye = torch.eye(m, dtype=torch.float32, device=mesh_in.device)
        batched = mesh_in.dim() == 3
        eye = eye.unsqueeze(0).expand(mesh_in.shape[0], m, m).contiguous() if batched else eye.unsqueeze(0)
        flat = position_attention(mesh_out, mesh_in, eye, head_scale(scale), locality, variant=self._variant)
        att = flat.reshape(flat.shape[0], flat.shape[1], -1, m).transpose(1, 2)  # (B|1, H, N, M)
        return att if batched else att[0]

    def convolution(self, A, U):
        eq = "bhnj,bjd->bnhd" if A.dim() == 4 else "hnj,bjd->bnhd"
        return torch.einsum(eq, A, U).reshape(U.shape[0], -1, self.n_head * U.shape[-1])


class posatt_cross(posatt):
    """Cross stage over per-sample meshes (pit.py:59-71)."""
    _cross = True


class posatt_fixed(posatt):
    """Self stage, one mesh shared by the batch (pit.py:129-144)."""


class posatt_cross_fixed(posatt_fixed):
    """Cross stage, shared meshes (pit.py:146-159)."""
    _cross = True


class posatt_periodic1d(posatt_fixed):
    """Shared 1-D mesh with periodic distance (pit.py:186-200)."""
    _variant = "periodic1d"


class posatt_cross_periodic1d(posatt_periodic1d):
    _cross = True


class posatt_periodic2d(posatt_fixed):
    """Shared 2-D mesh with periodic distance (pit.py:243-258)."""
    _variant = "periodic2d"


class posatt_cross_periodic2d(posatt_periodic2d):
    _cross = True


_FUSABLE_SELF = (posatt_fixed, posatt_periodic1d, posatt_periodic2d)
_FUSED_PROCESSOR = True


def use_fused_processor(enabled: bool) -> None:
    """Switch the fused processor (all blocks in one launch per direction) on or off; off runs attention and MLP per block."""
    global _FUSED_PROCESSOR
    _FUSED_PROCESSOR = bool(enabled)


_FUSABLE_CROSS = (posatt_cross_fixed, posatt_cross_periodic1d, posatt_cross_periodic2d)
_FUSED_DECODER_TAIL = True


def use_fused_decoder_tail(enabled: bool) -> None:
    """Switch the fused decoder tail (attention + MLP in one kernel) on or off; off runs `de(up(...))` as two ops."""
    global _FUSED_DECODER_TAIL
    _FUSED_DECODER_TAIL = bool(enabled)


class pit(nn.Module):
    """Encoder / processor / decoder container (pit.py:73-127).  Scripts subclass it and add ``forward``."""

    _cross_layer = posatt_cross
    _self_layer = posatt

    def __init__(self, space_dim, in_dim, out_dim, hid_dim, n_head, n_blocks, mesh_ltt, en_loc, de_loc):
        super().__init__()
        self.space_dim = space_dim
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.hid_dim = hid_dim
        self.n_head = n_head
        self.n_blocks = n_blocks
        self.mesh_ltt = None if mesh_ltt is None else mesh_ltt.reshape(-1, space_dim)
        self.en_local = en_loc
        self.de_local = de_loc

        # The reference builds every stage with the per-sample classes first and lets the
        # subclasses replace them (pit.py:182-184, 238-240, 296-298).  Drawing lmda in the same
        # order keeps seeded initialisations identical.
        self._build_attention(posatt_cross, posatt, first=True)
        if (self._cross_layer, self._self_layer) != (posatt_cross, posatt):
            self._build_attention(self._cross_layer, self._self_layer, first=False)

    def _build_attention(self, cross_cls, self_cls, first):
        h, hid = self.n_head, self.hid_dim
        self.down = cross_cls(h, self.in_dim, self.en_local)
        if first:
            self.en_layer = kaiming_mlp(h * (self.in_dim + self.space_dim), hid, hid)
        self.conv = nn.ModuleList([self_cls(h, hid, 1.0) for _ in range(self.n_blocks)])
        if first:
            self.mlp = nn.ModuleList([kaiming_mlp((1 + h) * hid, hid, hid) for _ in range(self.n_blocks)])
        self.up = cross_cls(h, hid, self.de_local)
        if first:
            self.de = kaiming_mlp(h * hid, hid, self.out_dim)

    @staticmethod
    def _mlp_gelu(mlp, x):
        return mlp.forward_gelu(x) if type(mlp) is kaiming_mlp else gelu(mlp(x))

    def encoder(self, mesh_in, func_in, mesh_ltt):
        return self._mlp_gelu(self.en_layer, self.down(mesh_ltt, mesh_in, func_in))

    def _fusable_processor(self, func_ltt, mesh_ltt):
        """Stock layers on a shared latent mesh, shapes the fused kernel covers (anything customised runs block by block)."""
        if not (_FUSED_PROCESSOR and len(self.conv) == len(self.mlp) > 0 and torch.is_tensor(mesh_ltt) and mesh_ltt.dim() == 2
                and torch.is_tensor(func_ltt) and func_ltt.is_cuda and func_ltt.dim() == 3):
            return False
        kind, hid, h = type(self.conv[0]), func_ltt.shape[-1], self.conv[0].n_head
        for attend, mix in zip(self.conv, self.mlp):
            if type(attend) is not kind or kind not in _FUSABLE_SELF or attend.n_head != h or attend.locality < 1.0:
                return False
            if (type(mix) is not kaiming_mlp or type(mix.mlp1) is not nn.Linear or type(mix.mlp2) is not nn.Linear
                    or mix.mlp1.bias is None or mix.mlp2.bias is None or tuple(mix.mlp1.weight.shape) != (hid, (1 + h) * hid)
                    or tuple(mix.mlp2.weight.shape) != (hid, hid) or mix.mlp1.weight.dtype != torch.float32):
                return False
        return processor_supported(mesh_ltt, func_ltt, h, len(self.conv), kind._variant)

    def processor(self, func_ltt, mesh_ltt):
        if self._fusable_processor(func_ltt, mesh_ltt):
            scales = head_scale(torch.stack([attend.lmda.reshape(-1) for attend in self.conv]))
            weights = [w for mix in self.mlp for w in (mix.mlp1.weight, mix.mlp1.bias, mix.mlp2.weight, mix.mlp2.bias)]
            return processor_blocks(func_ltt, mesh_ltt, scales, self.conv[0].n_head, self.conv[0]._variant, weights)
        for attend, mix in zip(self.conv, self.mlp):
            func_ltt = self._mlp_gelu(mix, attend(mesh_ltt, func_ltt))
        return func_ltt

    def decoder(self, mesh_ltt, func_ltt, mesh_out):
        up, de = self.up, self.de
        # Fused tail: attention + both Linears + GELU in one kernel, nothing N x (H*D) wide is written to memory.
        # Only taken for the stock layer types on a shared mesh; anything customised runs the two modules as written.
        if (_FUSED_DECODER_TAIL and type(up) in _FUSABLE_CROSS and type(de) is kaiming_mlp and mesh_ltt.dim() == 2
                and not meshes_need_grad(mesh_out, mesh_ltt)
                and func_ltt.is_cuda and func_ltt.dim() == 3 and de.mlp1.in_features == up.n_head * func_ltt.shape[-1]
                and decoder_tail_supported(mesh_ltt, func_ltt, up.n_head, de.mlp1.out_features, de.mlp2.out_features)):
            return decoder_tail(mesh_out, mesh_ltt, func_ltt, head_scale(up.lmda), up.locality, de.mlp1.weight, de.mlp1.bias,
                                de.mlp2.weight, de.mlp2.bias, variant=up._variant)
        return de(up(mesh_out, mesh_ltt, func_ltt))


class pit_fixed(pit):
    _cross_layer = posatt_cross_fixed
    _self_layer = posatt_fixed


class pit_periodic1d(pit):
    _cross_layer = posatt_cross_periodic1d
    _self_layer = posatt_periodic1d


class pit_periodic2d(pit):
    _cross_layer = posatt_cross_periodic2d
    _self_layer = posatt_periodic2d
