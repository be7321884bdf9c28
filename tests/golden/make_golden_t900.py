"""Real-dimension golden fixture at the BENCHMARKED shape from the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_t900.py

`generated_training` (G = 10, M = 288, D = 1024, A = 128), zoneout cells, B = 10, L = 180, T = 900, tf = 1.0, train mode, the
reference's default initialisation (`torch.manual_seed(0); Tacotron()`).  The reference's `Tacotron.forward` + `TacotronLoss` +
`backward` run on the CPU (fp32, the reference's own arithmetic) with `F.dropout` / `torch.rand` wrapped so that every mask comes
from a seeded `torch.rand(shape)` stream that tests/t900_case.py can replay on the GPU box (the reference code is not edited).
Stored: outputs (pre, post, stop, encoder output, alignment top-2 + every 30th alignment row), loss terms, gradients (full for
tensors <= 200 k elements, norm / sum / max / strided sample for the large ones) and checksums of the regenerated weights, inputs
and masks.  ~8 MB compressed.
"""
import os
import sys
import json
import time
import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    sys.path.insert(0, REF)
    import torch
    import torch.nn.functional as F
    import utils  # noqa: F401  (must precede modules.tacotron2: circular import in the reference)
    from params.params import Params as hp
    from modules.tacotron2 import Tacotron, TacotronLoss
    import t900_case as C

    hp.load(os.path.join(REF, 'params', C.CONFIG + '.json'))
    hp.load_state_dict(C.HP_OVERRIDES)
    hp.language_number = len(hp.languages) if hp.multi_language else 0
    hp.speaker_number = 0
    G = hp.language_number
    torch.manual_seed(C.WEIGHT_SEED)
    model = Tacotron().train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # our host classes must construct the very same parameters from the same seed (the fixture stores no weights)
    from multilingual_text_to_speech_b200 import configs
    from multilingual_text_to_speech_b200.modules.tacotron2 import Tacotron as OwnTacotron
    configs.apply(C.CONFIG, **C.HP_OVERRIDES)
    torch.manual_seed(C.WEIGHT_SEED)
    own = OwnTacotron().state_dict()
    assert list(own.keys()) == list(sd0.keys()) and all(torch.equal(own[k], sd0[k]) for k in sd0), 'seeded init differs'

    inp = C.build_inputs(hp.symbols_count() + 3, hp.num_mels, hp.stop_frames, G)
    plan = C.draw_plan(G, hp.prenet_dimension, hp.decoder_dimension, hp.num_mels, hp.postnet_dimension, hp.postnet_blocks)
    real_dropout, real_rand = F.dropout, torch.rand
    record = []

    def taped_dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        r = real_rand(tuple(input.shape))
        record.append(('mask', float(p), tuple(input.shape)))
        return input * (r >= p).to(input.dtype) * (1.0 / (1.0 - p))

    def taped_rand(*a, **k):
        r = real_rand(*a, **k)
        record.append(('rand', None, tuple(r.shape)))
        return r

    torch.manual_seed(C.TAPE_SEED)
    F.dropout, torch.rand = taped_dropout, taped_rand
    t0 = time.time()
    try:
        post, pre, stop, align, spk_pred, enc = model(inp['text'], inp['text_length'], inp['target'], inp['target_length'], None,
                                                      inp['languages'], 1.0)
    finally:
        F.dropout, torch.rand = real_dropout, real_rand
    print(f'reference forward: {time.time() - t0:.1f} s, {len(record)} draws')
    assert len(record) == len(plan), (len(record), len(plan))
    for (kind, p, shape), (name, pkind, pp, pshape) in zip(record, plan):
        assert kind == pkind and tuple(shape) == tuple(pshape) and (p is None or abs(p - pp) < 1e-9), (name, kind, p, shape, pkind, pp, pshape)

    crit = TacotronLoss(hp.guided_attention_steps, hp.guided_attention_toleration, hp.guided_attention_gain)
    loss, parts = crit(inp['text_length'], inp['target_length'], pre, inp['target'], post, inp['target'], stop, inp['stop_target'],
                       align, None, spk_pred, enc, None)
    t0 = time.time()
    loss.backward()
    print(f'reference backward: {time.time() - t0:.1f} s')
    losses = {k: float(v) for k, v in parts.items()}
    losses['total'] = float(loss)

    tape = C.replay_tape(G, hp.prenet_dimension, hp.decoder_dimension, hp.num_mels, hp.postnet_dimension, hp.postnet_blocks)
    out = {}
    top2 = torch.topk(align.detach(), 2, dim=2)
    out['out.pre'] = pre.detach().numpy()
    out['out.post'] = post.detach().numpy()
    out['out.stop'] = stop.detach().numpy()
    out['out.enc'] = enc.detach().numpy()
    out['out.align_top2_idx'] = top2.indices.numpy().astype(np.int16)
    out['out.align_top2_val'] = top2.values.numpy()
    out['out.align_rows'] = align.detach()[:, ::30].contiguous().numpy()          # every 30th decoder step, full rows
    out['out.align_rowsum'] = align.detach().sum(2).numpy()
    gmeta = {}
    for k, prm in model.named_parameters():
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        summ, sample = C.grad_summary(g)
        out['gsum.' + k] = summ
        if g.numel() <= C.FULL_GRAD_LIMIT:
            out['grad.' + k] = g.detach().numpy()
        else:
            out['gsample.' + k] = sample
        gmeta[k] = int(g.numel())
    for k, v in model.state_dict().items():
        if 'running_' in k or 'num_batches' in k:
            out['sd_after.' + k] = v.detach().numpy()
    meta = dict(config=C.CONFIG, hp_overrides=C.HP_OVERRIDES, B=C.B, L=C.L, T=C.T, tf=1.0, train=True, losses=losses,
                guided_g=hp.guided_attention_toleration, weight_digest=C.state_digest(sd0),
                input_digest={k: C.digest(v) for k, v in inp.items()}, tape_digest=C.tape_digest(tape), grad_numel=gmeta,
                torch=torch.__version__, noise=dict())
    out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, C.NAME + '.npz')
    np.savez_compressed(path, **out)
    print(f'{C.NAME}: {os.path.getsize(path) / 2 ** 20:.1f} MiB, losses={losses}')


if __name__ == '__main__':
    main()
