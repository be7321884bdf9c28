"""Time the fused decoder (forward + backward) alone at a given shape; prints per-phase CUDA-event times."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=64)
    ap.add_argument('--L', type=int, default=180)
    ap.add_argument('--T', type=int, default=900)
    ap.add_argument('--M', type=int, default=288)
    ap.add_argument('--kind', default='dropout')
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--precision', default='fp32')
    a = ap.parse_args()
    import decoder_cases as dc
    from multilingual_text_to_speech_b200 import functional as F, _lib
    c = dc.full_dim_case(B=a.B, L=a.L, T=a.T, M=a.M, kind=a.kind, seed=1, ragged=False)
    dev = torch.device('cuda:0')
    cfg, params, memory = dc._cuda_inputs(c, dev)
    target, lens = c.target.to(dev), c.lengths.to(dev)
    _lib.set_precision(a.precision)
    for it in range(a.iters + 1):
        n0 = _lib.launch_count()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        spec, stop, align = F.decoder_forward(cfg, memory, target, lens, params)
        e1.record()
        if not a.fwd_only:
            loss = spec.sum() + stop.sum() + (align * align).sum()
            loss.backward()
        e2.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f'iter {it}: fwd {e0.elapsed_time(e1):.2f} ms  bwd {e1.elapsed_time(e2):.2f} ms  host-enqueue {1e3 * (t1 - t0):.1f} ms '
              f'wall {1e3 * (t2 - t0):.1f} ms  launches {_lib.launch_count() - n0}  '
              f'frames/s {a.B * a.T / (t2 - t0):.0f}', flush=True)
        if a.precision == 'bf16' and it == a.iters:
            import ctypes
            F.PROFILE['keep_ws'] = True
            spec, stop, align = F.decoder_forward(cfg, memory, target, lens, params)
            if not a.fwd_only:
                (spec.sum() + stop.sum() + (align * align).sum()).backward()
            torch.cuda.synchronize()
            if not a.fwd_only:
                shp = F.PROFILE['last_shape']
                bnames = [['cell', 'barrier1', 'tma+tcgen05', 'tmem+store', 'barrier2', '-', '-', '-'],
                          ['attn:dw/softmax', 'attn:mma+dcum', 'barrier1', 'cell', 'barrier2', 'product', 'barrier3', '-']]
                for which, loop in enumerate(('gen-bwd', 'att-bwd')):
                    boff = _lib.load().b200tts_debug_persist_bwd_profile_offset(ctypes.byref(shp), which)
                    ns = 8
                    rb = F.PROFILE['last_bws'][boff:boff + 148 * ns * 8].view(torch.int64).view(148, ns).cpu().double()
                    act = rb[(rb.sum(1) > 0) & (rb.abs().max(1).values < 1e12)]        # rows of CTAs that do not exist hold whatever the allocator left there
                    if len(act):
                        print(f'{loop} loop: cycles/step by phase, CTA0 | mean | max over CTAs')
                        for j, nme in enumerate(bnames[which]):
                            if nme != '-':
                                print(f'   {nme:18s} {rb[0, j] / a.T:9.0f} {act[:, j].mean() / a.T:9.0f} {act[:, j].max() / a.T:9.0f}')
            off = _lib.load().b200tts_debug_persist_profile_offset(ctypes.byref(F.PROFILE['last_shape']))
            raw4 = F.PROFILE['last_ws'][off:off + 4 * 148 * 8 * 8].view(torch.int64).view(4, 148, 8).cpu().double()
            raw = raw4[:2]
            names = ['gemm+tmem', 'cell+q', 'barrier1', 'attn:q/load', 'attn:energy', 'attn:softmax', 'attn:ctx', 'barrier2']
            for k, loop in enumerate(('att', 'gen')):
                act = raw[k][(raw[k].sum(1) > 0) & (raw[k].abs().max(1).values < 1e12)]
                if len(act):
                    print(f'{loop} loop: cycles/step by phase, CTA0 | mean | max over CTAs')
                    for j, nme in enumerate(names):
                        print(f'   {nme:18s} {raw[k][0, j] / a.T:9.0f} {act[:, j].mean() / a.T:9.0f} {act[:, j].max() / a.T:9.0f}')
        if a.precision == 'bf16' and it == a.iters:
            rn = ['mma ctx: wait 1st box', 'mma ctx: rest', 'mma h: wait 1st box', 'mma h: rest', 'tma ctx: proxy fence', 'tma ctx: issue',
                  'tma h: proxy fence', 'tma h: issue']
            for k, loop in enumerate(('att', 'gen')):
                rr = raw4[2 + k]
                act = rr[(rr.sum(1) > 0) & (rr.abs().max(1).values < 1e12)]
                if len(act):
                    print(f'{loop} loop role threads: cycles/step, CTA0 | mean | max')
                    for j, nme in enumerate(rn):
                        print(f'   {nme:22s} {rr[0, j] / a.T:9.0f} {act[:, j].mean() / a.T:9.0f} {act[:, j].max() / a.T:9.0f}')
        for p in params:
            p.grad = None
        memory.grad = None


if __name__ == '__main__':
    main()
