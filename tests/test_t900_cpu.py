"""The real-dimension T = 900 fixture of the unmodified reference (tests/golden/t900_generated_training.npz):
(1) the seeded weights / inputs / mask tape regenerate bit-identically here, (2) the CPU oracle reproduces the reference's
decoder outputs at the real dimensions (D = 1024, A = 128, M = 288, G = 10) on a prefix of the frames."""
import types
import torch

import t900_case as C
from helpers import assert_close
from oracle import tacotron_oracle as O


def test_regenerated_state_matches_fixture():
    fx = C.Fixture()
    hp, model, inp = C.configure()
    assert C.state_digest(model.state_dict()) == fx.meta['weight_digest']
    assert {k: C.digest(v) for k, v in inp.items()} == fx.meta['input_digest']
    assert C.tape_digest(C.tape_for(hp)) == fx.meta['tape_digest']
    assert fx.out['pre'].shape == (C.B, hp.num_mels, C.T) and fx.out['align_top2_idx'].shape == (C.B, C.T, 2)


def test_oracle_matches_reference_on_a_prefix_at_real_dimensions():
    """Teacher-forced decoder outputs of frame t depend on targets < t only, so the oracle run on the first F frames must
    reproduce the reference's first F frames of `pre` / stop / alignment and the whole encoder output."""
    F_ = 24
    fx = C.Fixture()
    hp, model, inp = C.configure()
    tape = {k: (v if k == 'teacher' else v.float()) for k, v in C.tape_for(hp, frames=F_).items()}
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    for k in list(sd):
        if k.startswith('_decoder._prenet.') or k.startswith('_decoder._attention.'):
            sd[k] = sd[k[len('_decoder.'):]]
    ns = types.SimpleNamespace(**hp.state_dict())
    tl = torch.clamp(inp['target_length'], max=F_)
    with torch.no_grad():
        post, pre, stop, align, _, enc = O.tacotron_forward(sd, ns, inp['text'], inp['text_length'], inp['target'][:, :, :F_], tl, None,
                                                            inp['languages'], tape, training=True)
    assert_close(enc, fx.out['enc'], 1e-3, 1e-5, 'enc')
    assert_close(pre, fx.out['pre'][:, :, :F_], 1e-3, 1e-5, 'pre')
    real = fx.out['stop'][:, :F_] < 999.0
    assert_close(stop[real], fx.out['stop'][:, :F_][real], 1e-3, 1e-5, 'stop')
    assert_close(align[:, 0], fx.out['align_rows'][:, 0], 1e-3, 1e-7, 'alignment of step 0')
    idx, val = fx.out['align_top2_idx'][:, :F_].long(), fx.out['align_top2_val'][:, :F_]
    clear = (val[..., 0] - val[..., 1]) > 1e-7
    assert torch.equal(align.argmax(2)[clear], idx[..., 0][clear])
    assert_close(align.max(2).values, val[..., 0], 1e-3, 1e-7, 'alignment maxima')
