"""Data-parallel gradient exchange: one flat fp32 gradient bucket per model, one all-reduce per step.

Replaces torch.nn.DataParallel's per-step broadcast / gather / reduce-add of the reference (train.py:173-179, 255-256;
SURVEY.md sections 2.2, 8e) with one process per GPU and a single NCCL all-reduce over NVLink of a flat buffer that
every parameter's .grad is a view of (the backward kernels accumulate straight into it -- no packing copy).
Utterances shard naturally: rank r processes its own batch, BatchNorm statistics stay per rank (DataParallel semantics).
"""
import torch
import torch.distributed as dist

FLAT_ALIGN = 4      # floats: every tensor starts on a 16-byte boundary (vectorised / TMA paths of the kernels need it)


def flat_layout(params):
    """Offsets of the tensors inside a flat buffer, each rounded up to FLAT_ALIGN floats; shared by GradBucket, FlatParams and the
    Adam moment buffers so that element i of every buffer belongs to the same parameter element.  -> (offsets, total)."""
    offsets, off = [], 0
    for p in params:
        offsets.append(off)
        off += (p.numel() + FLAT_ALIGN - 1) // FLAT_ALIGN * FLAT_ALIGN
    return offsets, off


class GradBucket:
    def __init__(self, model, world_size=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.offsets, total = flat_layout(self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.bind()

    def _view(self, k):
        p = self.params[k]
        return self.flat[self.offsets[k]:self.offsets[k] + p.numel()].view_as(p)

    def bind(self):
        """(Re-)attach every parameter's .grad to its view of the flat buffer.  A gradient found OUTSIDE the buffer (someone called
        `zero_grad(set_to_none=True)` and autograd allocated a fresh tensor) is added in first, so no gradient is ever dropped."""
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * self.flat.element_size()
        for k, p in enumerate(self.params):
            view = self._view(k)
            g = p.grad
            if g is not None and not (lo <= g.data_ptr() < hi):
                view.add_(g.to(view.dtype))
            if g is None or g.data_ptr() != view.data_ptr():
                p.grad = view          # autograd accumulates in place into the view

    def zero(self):
        self.bind()
        self.flat.zero_()

    def allreduce(self):
        """Mean over ranks of the per-rank mean-loss gradients (== DataParallel's gradient of the global-batch mean)."""
        self.bind()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)
        return self.flat

    def grad_norm(self):
        return torch.linalg.vector_norm(self.flat)
