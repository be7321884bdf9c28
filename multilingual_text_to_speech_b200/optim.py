"""Optimizer step of the training loop (reference train.py:84-85, 260-271) on flat buffers.

The reference clips the global gradient norm (`clip_grad_norm_(model.parameters(), hp.gradient_clipping)`) and steps
`torch.optim.Adam(lr, weight_decay)` (coupled L2 decay, not AdamW) over ~400 parameter tensors with a `StepLR` schedule.  Here every
parameter is a view of ONE flat fp32 buffer (`FlatParams`), its gradient a view of the all-reduced flat gradient bucket
(`distributed.GradBucket`), and one library call (`b200tts_adam_clip_step`: norm, clip, Adam, three launches) updates the model.

Difference to torch.optim.Adam worth knowing: a parameter that received NO gradient in a step has a zero (not a None) gradient here,
so it is still weight-decayed and its moments decay -- torch skips such parameters.  Every parameter of the Tacotron model receives a
gradient in every training step, so the two coincide on the hot path (tests/test_gpu_optim.py checks the real model).
"""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr
from .distributed import flat_layout


class FlatParams:
    """Re-homes the trainable parameters of `model` as views of one flat buffer (same order and same 16-byte aligned layout as
    `GradBucket`; the padding elements stay zero in the parameter, gradient and moment buffers)."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        ref = self.params[0]
        self.offsets, total = flat_layout(self.params)
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view


class FusedAdam:
    """clip_grad_norm_ + Adam(weight_decay) + StepLR on the flat buffers; state layout and hyper-parameter names follow torch.optim."""

    def __init__(self, flat_params, grad_bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None,
                 lr_decay_every=None, lr_decay=1.0):
        assert flat_params.flat.numel() == grad_bucket.flat.numel() and flat_params.offsets == grad_bucket.offsets, \
            'parameter and gradient buffers differ in layout'
        self.bucket = grad_bucket
        self.p, self.g = flat_params.flat, grad_bucket.flat
        self.m, self.v = torch.zeros_like(self.p), torch.zeros_like(self.p)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self.lr_decay_every, self.lr_decay = lr_decay_every, float(lr_decay)
        self.steps = 0
        self._scratch = None

    def current_lr(self):
        """StepLR(step_size=lr_decay_every, gamma=lr_decay) evaluated at the number of steps taken (train.py:270)."""
        if not self.lr_decay_every:
            return self.lr
        return self.lr * self.lr_decay ** (self.steps // self.lr_decay_every)

    def step(self):
        """One update.  Returns a 2-element device tensor: (gradient norm before clipping, applied clip coefficient)."""
        if not self.p.is_cuda:
            raise _lib.B200TTSError('FusedAdam needs CUDA buffers (there is no CPU fallback)')
        lib = _lib.load()
        self.bucket.bind()          # a gradient that escaped the bucket (zero_grad(set_to_none=True)) is folded back in, never dropped
        if self._scratch is None:
            self._scratch = torch.zeros(lib.b200tts_adam_clip_scratch_floats(), dtype=torch.float32, device=self.p.device)
        lr = self.current_lr()
        self.steps += 1
        check(lib.b200tts_adam_clip_step(ptr(self.p), ptr(self.g), ptr(self.m), ptr(self.v), self.p.numel(), lr, self.betas[0], self.betas[1],
                                         self.eps, self.weight_decay, self.max_grad_norm if self.max_grad_norm else 0.0, self.steps,
                                         ptr(self._scratch), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              'b200tts_adam_clip_step')
        return self._scratch[:2]

    def state_dict(self):
        return {'steps': self.steps, 'exp_avg': self.m, 'exp_avg_sq': self.v, 'lr': self.lr, 'betas': self.betas, 'eps': self.eps,
                'weight_decay': self.weight_decay, 'max_grad_norm': self.max_grad_norm, 'lr_decay_every': self.lr_decay_every,
                'lr_decay': self.lr_decay}

    def load_state_dict(self, d):
        self.steps = int(d['steps'])
        self.m.copy_(d['exp_avg']); self.v.copy_(d['exp_avg_sq'])
        for k in ('lr', 'betas', 'eps', 'weight_decay', 'max_grad_norm', 'lr_decay_every', 'lr_decay'):
            if k in d:
                setattr(self, k, d[k])
