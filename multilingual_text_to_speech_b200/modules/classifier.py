"""Adversarial speaker classifier with gradient reversal (surface of reference modules/classifier.py:6-69)."""
import torch
from torch.nn import Sequential, Linear

from .. import functional as F


class GradientReversalFunction(torch.autograd.Function):
    """Identity forward; backward clamps to [-c, c] and multiplies by -l (classifier.py:6-18)."""

    @staticmethod
    def forward(ctx, x, l, c):
        ctx.l, ctx.c = l, c
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return -ctx.l * grad_output.clamp(-ctx.c, ctx.c), None, None


class ReversalClassifier(torch.nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, gradient_clipping_bounds, scale_factor=1.0):
        super().__init__()
        self._lambda = scale_factor
        self._clipping = gradient_clipping_bounds
        self._output_dim = output_dim
        self._classifier = Sequential(Linear(input_dim, hidden_dim), Linear(hidden_dim, output_dim))

    def forward(self, x):
        x = GradientReversalFunction.apply(x, self._lambda, self._clipping)
        for layer in self._classifier:
            x = F.linear(x, layer.weight, layer.bias)
        return x

    @staticmethod
    def loss(input_lengths, speakers, prediction, embeddings=None):
        """Masked cross entropy over valid input positions (classifier.py:60-69)."""
        # the reference sizes the mask by max(input_lengths), which must equal the padded length of `prediction` for its cross entropy
        # to be shape-valid; using that padded length directly avoids a device -> host read (the step stays CUDA-graph capturable)
        ml = prediction.shape[1]
        mask = torch.arange(ml, device=prediction.device)[None, :] < input_lengths.to(prediction.device)[:, None]
        target = torch.where(mask, speakers[:, None].expand(-1, ml), torch.full_like(mask, -100, dtype=speakers.dtype))
        return torch.nn.functional.cross_entropy(prediction.transpose(1, 2), target, ignore_index=-100)
