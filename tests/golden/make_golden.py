"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the reference's modules read-only from /root/reference (recipe: SURVEY.md appendix C), builds
small-dimension models for every covered configuration family, runs ``Tacotron.forward`` +
``TacotronLoss`` + ``backward`` with a *recorded dropout-mask tape* (``torch.nn.functional.dropout`` and
``torch.rand`` are wrapped so every mask / teacher-forcing coin the reference draws is captured in call
order -- the reference code itself is not edited), and stores inputs, weights, masks, outputs, loss
terms and all parameter gradients into ``tests/golden/<case>.npz``.

The fixtures pin ``oracle/tacotron_oracle.py`` (tests/test_oracle_golden.py) and, on the GPU box,
the CUDA path (tests/test_gpu_golden.py).  /root/reference does not exist on the GPU box, which
is why the vectors are committed.
"""
import os
import sys
import json
import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

SMALL = dict(embedding_dimension=32, encoder_dimension=32, prenet_dimension=24, attention_dimension=16,
             attention_kernel_size=7, attention_location_dimension=8, decoder_dimension=48, postnet_dimension=32,
             num_mels=12, characters='abcdefgh', punctuations_out='.,', punctuations_in="'",
             reversal_classifier_dim=16, speaker_embedding_dimension=8)

CASES = {
    # name: (hp overrides, B, L, T, teacher_forcing, train_mode)
    'lj_dropout': (dict(), 5, 11, 17, 1.0, True),
    'lj_zoneout': (dict(decoder_regularization='zoneout'), 5, 11, 17, 1.0, True),
    'lj_mixed_tf': (dict(), 4, 9, 13, 0.5, True),
    'lj_eval_free': (dict(), 3, 9, 10, 0.0, False),
    'generated_training': (dict(encoder_type='generated', multi_language=True, languages=['a', 'b', 'c'],
                                language_embedding_dimension=6, generator_dim=5, generator_bottleneck_dim=3),
                           6, 11, 17, 1.0, True),
    'shared_switching': (dict(encoder_type='simple', multi_language=True, multi_speaker=True,
                              languages=['a', 'b', 'c'], language_embedding_dimension=4,
                              reversal_classifier=True, reversal_classifier_w=0.5), 6, 11, 17, 1.0, True),
    'generated_switching': (dict(encoder_type='generated', multi_language=True, multi_speaker=True,
                                 languages=['a', 'b'], language_embedding_dimension=0, generator_dim=4,
                                 generator_bottleneck_dim=2, reversal_classifier=True,
                                 reversal_classifier_w=0.125, decoder_regularization='zoneout'),
                            6, 12, 15, 1.0, True),
}


def main():
    sys.path.insert(0, REF)
    import torch
    import torch.nn.functional as F
    import utils  # noqa: F401  (must precede modules.tacotron2: circular import in the reference)
    from params.params import Params as hp
    from modules.tacotron2 import Tacotron, TacotronLoss

    defaults = dict(hp.state_dict())
    real_dropout, real_rand = F.dropout, torch.rand
    record = {'masks': [], 'rands': []}

    def taped_dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        keep = (torch.rand_like(input) >= p).to(input.dtype)
        record['masks'].append((float(p), keep.detach().clone()))
        return input * keep * (1.0 / (1.0 - p))

    def taped_rand(*a, **k):
        r = real_rand(*a, **k)
        record['rands'].append(r.detach().clone())
        return r

    for name, (over, B, L, T, tf, train_mode) in CASES.items():
        hp.load_state_dict(defaults)
        hp.load_state_dict(SMALL)
        hp.load_state_dict(over)
        hp.language_number = len(hp.languages) if hp.multi_language else 0
        hp.speaker_number = 3 if hp.multi_speaker else 0
        G = max(hp.language_number, 1)
        torch.manual_seed(sum(map(ord, name)))
        model = Tacotron()
        # sharpen: trained-like attention / gates so argmax parity is meaningful
        with torch.no_grad():
            model._attention._energy.weight.mul_(6.0)
            model._attention._memory.weight.mul_(3.0)
            model._attention._query.weight.mul_(3.0)
            for prm in model._decoder._attention_lstm.parameters():
                prm.mul_(2.0)
        model.train(train_mode)
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        nsym = hp.symbols_count() + 3
        lens = torch.sort(torch.randint(max(2, L // 2), L + 1, (B,)), descending=True).values
        lens[0] = L
        text = torch.randint(1, nsym, (B, L))
        for b in range(B):
            text[b, lens[b]:] = 0
        tlens = torch.clamp(lens * T // L + torch.randint(0, 2, (B,)), max=T)
        tlens[0] = T
        mel = torch.randn(B, hp.num_mels, T)
        spk = torch.randint(0, hp.speaker_number, (B,)) if hp.multi_speaker else None
        lang = (torch.arange(B) % G) if hp.multi_language else None
        stop_t = torch.zeros(B, T)
        for b in range(B):
            stop_t[b, tlens[b] - 2:] = 1.0

        record['masks'].clear(); record['rands'].clear()
        F.dropout, torch.rand = taped_dropout, taped_rand
        try:
            if train_mode:
                post, pre, stop, align, spk_pred, enc = model(text, lens, mel, tlens, spk, lang, tf)
            else:
                with torch.no_grad():
                    post, pre, stop, align, spk_pred, enc = model(text, lens, mel, tlens, spk, lang, tf)
        finally:
            F.dropout, torch.rand = real_dropout, real_rand

        out = {}
        grads = {}
        losses = {}
        if train_mode:
            crit = TacotronLoss(hp.guided_attention_steps, hp.guided_attention_toleration, hp.guided_attention_gain)
            loss, parts = crit(lens, tlens, pre, mel, post, mel, stop, stop_t, align, spk, spk_pred, enc,
                               model._reversal_classifier if hp.reversal_classifier else None)
            loss.backward()
            losses = {k: float(v) for k, v in parts.items()}
            losses['total'] = float(loss)
            seen = set()
            for k, prm in model.named_parameters():
                grads[k] = prm.grad.detach().numpy() if prm.grad is not None else np.zeros(tuple(prm.shape), np.float32)
                seen.add(k)

        # ---- assemble the named tape from the recorded call order ----
        masks = list(record['masks'])
        zone = hp.decoder_regularization == 'zoneout'
        tape = {}
        pos = 0
        if train_mode:
            n_enc = 14 if hp.encoder_type == 'generated' else hp.encoder_blocks
            for j in range(n_enc):
                tape[f'enc{j}'] = masks[pos][1]; pos += 1
        tape['prenet0'] = masks[pos][1]; pos += 1
        tape['prenet1'] = masks[pos][1]; pos += 1
        teacher = (record['rands'][0] > (1 - tf))
        tape['teacher'] = teacher
        D, P = hp.decoder_dimension, hp.prenet_dimension
        per = {k: torch.ones(T, B, D) for k in ('att_h', 'att_c', 'gen_h', 'gen_c')}
        sp0 = torch.ones(T, B, P); sp1 = torch.ones(T, B, P)
        for i in range(T):
            if not bool(teacher[i]):
                sp0[i] = masks[pos][1]; pos += 1
                sp1[i] = masks[pos][1]; pos += 1
            if train_mode:
                per['att_h'][i] = masks[pos][1]; pos += 1
                if zone:
                    per['att_c'][i] = masks[pos][1]; pos += 1
                per['gen_h'][i] = masks[pos][1]; pos += 1
                if zone:
                    per['gen_c'][i] = masks[pos][1]; pos += 1
        tape.update(per)
        tape['step_prenet0'], tape['step_prenet1'] = sp0, sp1
        if train_mode:
            for j in range(hp.postnet_blocks):
                tape[f'post{j}'] = masks[pos][1]; pos += 1
        assert pos == len(masks), (name, pos, len(masks))

        out['meta'] = np.frombuffer(json.dumps(dict(
            hp={k: v for k, v in hp.state_dict().items() if isinstance(v, (int, float, str, bool, list))},
            B=B, L=L, T=T, tf=tf, train=train_mode, losses=losses,
            guided_g=hp.guided_attention_toleration)).encode(), dtype=np.uint8)
        out['in.text'] = text.numpy(); out['in.text_length'] = lens.numpy()
        out['in.target'] = mel.numpy(); out['in.target_length'] = tlens.numpy()
        out['in.stop_target'] = stop_t.numpy()
        if spk is not None: out['in.speakers'] = spk.numpy()
        if lang is not None: out['in.languages'] = lang.numpy()
        for k, v in sd0.items():
            out['sd.' + k] = v.numpy()
        for k, v in model.state_dict().items():
            if 'running_' in k or 'num_batches' in k:
                out['sd_after.' + k] = v.detach().numpy()
        for k, v in tape.items():
            out['tape.' + k] = v.numpy().astype(np.uint8)
        for k, v in (('post', post), ('pre', pre), ('stop', stop), ('align', align), ('spk_pred', spk_pred), ('enc', enc)):
            if v is not None:
                out['out.' + k] = v.detach().numpy()
        for k, v in grads.items():
            out['grad.' + k] = v
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB, masks={len(masks)}, losses={losses}')


if __name__ == '__main__':
    main()
