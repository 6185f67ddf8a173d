"""Drop-in for the reference ``utils.py``: losses, parameter count and the pixel-wise normaliser.

Same names and call semantics (``RelLpNorm(out_dim, p)(true, pred)`` etc., utils.py:6-98); these sit
outside the position-attention hot path and are ordinary torch code -- except that ``RelLpNorm`` on CUDA float32
tensors with p in {1, 2} runs as three launches of libpit_posatt.so instead of ~14 torch kernels.
"""
import operator
from functools import reduce

import torch
import torch.nn.functional as F

# the reference module has no __all__, so `from utils import *` also hands out its imports (utils.py:1-4)
__all__ = ["PixelWiseNormalization", "count_params", "RelMaxNorm", "RelLpNorm", "operator", "reduce", "torch", "F"]


def count_params(model) -> int:
    """Number of scalar parameters (utils.py:52-57)."""
    return sum(p.numel() for p in model.parameters())


class _RelativeError:
    """Per-sample relative error, averaged over the output variables and SUMMED over the batch
    (utils.py:77, 98) -- which is why data-parallel gradients are reduced with SUM, not MEAN."""

    def __init__(self, out_dim):
        self._out_dim = out_dim

    def _magnitude(self, x):  # (batch, L, out_dim) -> (batch, out_dim)
        raise NotImplementedError

    def __call__(self, true, pred):
        t = true.reshape(true.size(0), -1, self._out_dim)
        q = pred.reshape(pred.size(0), -1, self._out_dim)
        return (self._magnitude(t - q) / self._magnitude(t)).mean(dim=-1).sum()


class RelLpNorm(_RelativeError):
    def __init__(self, out_dim, p):
        super().__init__(out_dim)
        self._ord = p

    def _magnitude(self, x):
        return torch.norm(x, p=self._ord, dim=1)

    def __call__(self, true, pred):
        t = true.reshape(true.size(0), -1, self._out_dim)
        q = pred.reshape(pred.size(0), -1, self._out_dim)
        if q.is_cuda:
            from .posatt import rel_lp_loss, rel_lp_supported       # the CUDA library is only needed for CUDA tensors
            if rel_lp_supported(t, q, self._ord):
                return rel_lp_loss(t, q, self._ord)
        return super().__call__(true, pred)


class RelMaxNorm(_RelativeError):
    def _magnitude(self, x):
        return x.abs().amax(dim=1)


class PixelWiseNormalization:
    """Per-pixel standardisation fitted on (n, h, w, c) data; statistics are bilinearly resampled when
    applied at another resolution (zero-shot super-resolution, utils.py:6-50)."""

    def __init__(self, x, eps=1e-5):
        self.mean = x.mean(dim=0, keepdim=True)
        self.std = x.std(dim=0, keepdim=True)
        self.eps = eps

    def _stats_for(self, x):
        if x.shape[1:] == self.mean.shape[1:] or x.dim() != 4:
            return self.mean, self.std
        size = (x.shape[1], x.shape[2])
        resample = lambda s: F.interpolate(s.permute(0, 3, 1, 2), size=size, mode="bilinear",
                                           align_corners=False).permute(0, 2, 3, 1)
        return resample(self.mean), resample(self.std)

    def normalize(self, x):
        mean, std = self._stats_for(x)
        return (x - mean) / (std + self.eps)

    def denormalize(self, x):
        mean, std = self._stats_for(x)
        return x * (std + self.eps) + mean

    def to(self, device):
        self.mean, self.std = self.mean.to(device), self.std.to(device)
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")
