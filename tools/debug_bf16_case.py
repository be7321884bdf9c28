import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import decoder_cases as dc
from multilingual_text_to_speech_b200 import functional as F, _lib


def grads(c, mode):
    dev = torch.device('cuda:0')
    cfg, params, memory = dc._cuda_inputs(c, dev)
    _lib.set_precision(mode)
    try:
        spec, stop, align = F.decoder_forward(cfg, memory, c.target.to(dev), c.lengths.to(dev), params)
        g = torch.Generator().manual_seed(99)
        rs = [torch.randn(t.shape, generator=g, dtype=torch.float64).float().to(dev) for t in (spec, stop, align)]
        ((spec * rs[0]).sum() + (stop * rs[1]).sum() + (align * rs[2]).sum()).backward()
        torch.cuda.synchronize()
    finally:
        _lib.set_precision('fp32')
    return {'memory': memory.grad.cpu(), **{f: p.grad.cpu() for (f, k), p in zip(dc.PARAM_KEYS, params)}}


for kw in [dict(B=40, L=64, T=24, kind='zoneout', seed=1), dict(B=40, L=64, T=24, kind='dropout', seed=1),
           dict(B=40, L=70, T=24, kind='zoneout', seed=1), dict(B=24, L=64, T=24, kind='zoneout', seed=1),
           dict(B=40, L=64, T=8, kind='zoneout', seed=1)]:
    c = dc.full_dim_case(**kw)
    g32, g16 = grads(c, 'fp32'), grads(c, 'bf16')
    worst = sorted(((float((g16[k] - g32[k]).norm() / (g32[k].norm() + 1e-12)), k) for k in g32), reverse=True)[:5]
    print(kw, [(k, f'{v:.3f}') for v, k in worst], 'lens', c.lengths.tolist()[:6], flush=True)
    if kw['T'] == 24 and kw['B'] == 40 and kw['L'] == 64 and kw['kind'] == 'zoneout':
        d = (g16['memory'] - g32['memory'])
        rel_b = d.flatten(1).norm(dim=1) / (g32['memory'].flatten(1).norm(dim=1) + 1e-12)
        print('  per-utterance rel err of d_memory:', [f'{x:.2f}' for x in rel_b.tolist()])
        rel_l = d.permute(1, 0, 2).flatten(1).norm(dim=1) / (g32['memory'].permute(1, 0, 2).flatten(1).norm(dim=1) + 1e-12)
        print('  per-position rel err:', [f'{x:.2f}' for x in rel_l.tolist()])
