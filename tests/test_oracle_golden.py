"""Pin the CPU oracle (oracle/tacotron_oracle.py) against golden vectors produced by the unmodified reference."""
import pytest
import torch

from helpers import Golden, GOLDEN_CASES, assert_close
from oracle import tacotron_oracle as O


def _run(g, dtype, with_grad):
    sd = g.cast_sd(dtype, requires_grad=with_grad)
    tape = g.tape_cast(dtype)
    i = g.inputs
    stats = {}
    out = O.tacotron_forward(sd, g.hp, i['text'], i['text_length'], i['target'].to(dtype), i['target_length'],
                             i.get('speakers'), i.get('languages'), tape, training=g.train, stats=stats)
    return sd, out, stats


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_forward_matches_reference(name):
    g = Golden(name)
    with torch.no_grad():
        _, (post, pre, stop, align, spk, enc), _ = _run(g, torch.float64, False)
    assert_close(enc, g.out['enc'], 1e-4, 1e-5, 'enc')
    assert_close(align, g.out['align'], 1e-4, 1e-5, 'align')
    assert_close(pre, g.out['pre'], 1e-4, 1e-5, 'pre')
    assert_close(stop, g.out['stop'], 1e-4, 1e-5, 'stop')
    assert_close(post, g.out['post'], 1e-3, 1e-4, 'post')
    if 'spk_pred' in g.out:
        assert_close(spk, g.out['spk_pred'], 1e-4, 1e-5, 'spk_pred')
    # bit-exact discrete decisions (north_star): alignment argmax and stop sign
    assert torch.equal(align.argmax(2), g.out['align'].argmax(2))
    assert torch.equal(stop > 0, g.out['stop'] > 0)


@pytest.mark.parametrize('name', [n for n in GOLDEN_CASES if n != 'lj_eval_free'])
def test_loss_and_gradients_match_reference(name):
    g = Golden(name)
    sd, (post, pre, stop, align, spk, enc), _ = _run(g, torch.float64, True)
    i = g.inputs
    tgt = i['target'].double()
    loss, parts = O.tacotron_loss(g.hp, g.meta['guided_g'], i['text_length'], i['target_length'], pre, tgt, post, tgt,
                                  stop, i['stop_target'], align, i.get('speakers'), spk)
    for k, v in parts.items():
        assert abs(float(v.detach()) - g.losses[k]) < 1e-4 * max(1.0, abs(g.losses[k])), (k, float(v.detach()), g.losses[k])
    loss.backward()
    for k, ref in g.grad.items():
        if k.startswith('_decoder._prenet.') or k.startswith('_decoder._attention.'):
            continue
        got = sd[k].grad
        got = torch.zeros_like(ref) if got is None else got.clone()
        if k == '_embedding.weight':
            got[0] = 0          # Embedding(padding_idx=0): row 0 receives no gradient (tacotron2.py:237-238)
        scale = float(ref.abs().max()) + 1e-12
        assert_close(got, ref, 2e-3, 2e-4 * scale + 1e-9, 'grad ' + k)


@pytest.mark.parametrize('name', ['lj_dropout', 'generated_training'])
def test_running_stats(name):
    g = Golden(name)
    with torch.no_grad():
        sd, _, stats = _run(g, torch.float64, False)
    for prefix, (mean, var, count) in stats.items():
        bn = prefix + ('._regularizer' if '_layers' in prefix else '._block.2')
        rm, rv = O.running_stats_update(sd[bn + '.running_mean'], sd[bn + '.running_var'], mean, var, count)
        assert_close(rm, g.sd_after[bn + '.running_mean'], 1e-4, 1e-6, bn + '.running_mean')
        assert_close(rv, g.sd_after[bn + '.running_var'], 1e-4, 1e-6, bn + '.running_var')
