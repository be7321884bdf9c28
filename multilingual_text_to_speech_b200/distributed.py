"""Data-parallel gradient exchange: one flat fp32 gradient bucket per model, one all-reduce per step.

Replaces torch.nn.DataParallel's per-step broadcast / gather / reduce-add of the reference (train.py:173-179, 255-256;
SURVEY.md sections 2.2, 8e) with one process per GPU and a single NCCL all-reduce over NVLink of a flat buffer that
every parameter's .grad is a view of (the backward kernels accumulate straight into it -- no packing copy).
Utterances shard naturally: rank r processes its own batch, BatchNorm statistics stay per rank (DataParallel semantics).
"""
import torch
import torch.distributed as dist


class GradBucket:
    def __init__(self, model, world_size=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)      # autograd accumulates in place into the view
            off += n

    def zero(self):
        self.flat.zero_()

    def allreduce(self):
        """Mean over ranks of the per-rank mean-loss gradients (== DataParallel's gradient of the global-batch mean)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)
        return self.flat

    def grad_norm(self):
        return torch.linalg.vector_norm(self.flat)
