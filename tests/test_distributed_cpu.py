"""world_size-2 gloo test of the data-parallel gradient exchange (host logic of the N > 1 path, runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multilingual_text_to_speech_b200.distributed import GradBucket
    torch.manual_seed(0)                       # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    bucket = GradBucket(model, world)
    g = torch.Generator().manual_seed(100 + rank)     # each rank owns a different shard of utterances
    x, y = torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)
    for _ in range(2):                         # two steps: zero() must really reset the shared buffer
        bucket.zero()
        torch.nn.functional.mse_loss(model(x), y).backward()
        flat = bucket.allreduce().clone()
    views_ok = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    out[rank] = (flat, x, y, views_ok)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_global_batch_gradient():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert torch.equal(res[0][0], res[1][0]), 'ranks disagree after the all-reduce'
    assert res[0][3] and res[1][3]
    # reference semantics: gradient of the mean loss over the GLOBAL batch (DataParallel gathers, then averages)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x = torch.cat([res[0][1], res[1][1]]); y = torch.cat([res[0][2], res[1][2]])
    torch.nn.functional.mse_loss(model(x), y).backward()
    from multilingual_text_to_speech_b200.distributed import flat_layout
    offsets, total = flat_layout(list(model.parameters()))          # tensors start on 16-byte boundaries; the padding stays zero
    ref = torch.zeros(total)
    for p, off in zip(model.parameters(), offsets):
        ref[off:off + p.numel()] = p.grad.flatten()
    assert torch.allclose(res[0][0], ref, rtol=1e-5, atol=1e-7)
