"""Per-language parameter generators (surface of reference modules/generated.py:7-96).

Parameter names (`_bottleneck`, `_kernel`, `_bias`, `_affine`, running_mean/var buffers) follow the reference so
checkpoints load strictly.  The arithmetic is the library's generator op; the convolution / batch-norm
themselves are fused into the conv-block op (see modules/layers.py), so these classes expose `generate`.
"""
import torch
from torch.nn import Linear

from .. import functional as F


class Conv1dGenerated(torch.nn.Module):
    """Generates one convolution kernel per group from a generator embedding (generated.py:7-42)."""

    def __init__(self, embedding_dim, bottleneck_dim, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias=True):
        super().__init__()
        assert stride == 1 and padding == 0, 'only stride 1 / explicit "same" padding is used by the encoders'
        self._in_channels, self._out_channels, self._kernel_size = in_channels, out_channels, kernel_size
        self._stride, self._padding, self._dilation, self._groups = stride, padding, dilation, groups
        self._bottleneck = Linear(embedding_dim, bottleneck_dim)
        self._kernel = Linear(bottleneck_dim, out_channels // groups * in_channels // groups * kernel_size)
        self._bias = Linear(bottleneck_dim, out_channels // groups) if bias else None

    def generate(self, generator_embedding):
        """-> kernel [out_channels, in_channels // groups, k] (the tensor F.conv1d received in the reference)."""
        assert generator_embedding.shape[0] == self._groups, \
            'Number of groups of a convolutional layer must match the number of generators.'
        flat = F.GeneratorFunction.apply(generator_embedding, self._bottleneck.weight, self._bottleneck.bias,
                                         self._kernel.weight, self._kernel.bias)
        return flat.view(self._out_channels, self._in_channels // self._groups, self._kernel_size)

    def forward(self, generator_embedding, x):
        """Standalone generated grouped convolution (generated.py:34-42): kernel synthesis + the conv-only stage of the block op.
        (Inside ConvBlockGenerated the convolution, batch norm and activation are ONE op.)"""
        kernel = self.generate(generator_embedding)
        out = F.conv_block(x, kernel, None, None, None, None, None, self._groups, self._kernel_size, self._dilation, 'identity', False,
                           self.training, 0.0, 0.0, 0.0, 0, stage=1)
        # the reference convolves WITHOUT padding (its caller pads, layers.py:112-115,126); the library op is 'same'-padded, and the
        # un-padded ("valid") result is exactly its centre
        pad = (self._kernel_size - 1) * self._dilation // 2
        if pad:
            out = out[:, :, pad:out.shape[2] - pad]
        if self._bias is not None:
            bias = F.GeneratorFunction.apply(generator_embedding, self._bottleneck.weight, self._bottleneck.bias, self._bias.weight,
                                             self._bias.bias)
            out = out + bias.reshape(1, -1, 1)
        return out


class BatchNorm1dGenerated(torch.nn.Module):
    """Batch normalisation whose affine parameters are generated per group (generated.py:45-96)."""

    def __init__(self, embedding_dim, bottleneck_dim, num_features, groups=1, eps=1e-8, momentum=0.1):
        super().__init__()
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        self._num_features = num_features // groups
        self._eps, self._momentum, self._groups = eps, momentum, groups
        self._bottleneck = Linear(embedding_dim, bottleneck_dim)
        self._affine = Linear(bottleneck_dim, self._num_features + self._num_features)

    def generate(self, generator_embedding):
        """-> affine [G, 2*C]: scale = [:, :C], bias = [:, C:]."""
        assert generator_embedding.shape[0] == self._groups, \
            'Number of groups of a batchnorm layer must match the number of generators.'
        return F.GeneratorFunction.apply(generator_embedding, self._bottleneck.weight, self._bottleneck.bias,
                                         self._affine.weight, self._affine.bias)

    def forward(self, generator_embedding, x):
        """Standalone generated batch norm (generated.py:71-96): affine synthesis + the batch-norm-only stage of the block op."""
        affine = self.generate(generator_embedding)
        C = self._num_features
        out = F.conv_block(x, None, affine[:, :C], affine[:, C:], self.running_mean, self.running_var, None, self._groups, 1, 1, 'identity',
                           False, self.training, self._eps, self._momentum, 0.0, 2 * C, stage=2)
        if self.training:
            self.num_batches_tracked += 1
        return out
