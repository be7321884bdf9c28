"""Tacotron.inference (synthesize.py entry point) against golden vectors of the UNMODIFIED reference's own `inference()`
(modules/tacotron2.py:387-408, :201-207; tests/golden/make_golden_inference.py): chunked free-running decode with carried state and
early exit, the always-on prenet dropout replayed from the recorded masks, and the per-character language-mixing branch of the
generated encoder (modules/encoder.py:213-219)."""
import json
import os
import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    tape = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('tape.')}
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('in.')}
    return meta, sd, tape, inp, torch.from_numpy(z['out.post'])


@pytest.mark.parametrize('name', ['inf_lj', 'inf_generated_mix'])
@pytest.mark.parametrize('chunk', [7, 128])
def test_inference_matches_reference(name, chunk):
    from multilingual_text_to_speech_b200 import functional as F
    from multilingual_text_to_speech_b200.params.params import Params as hp
    from multilingual_text_to_speech_b200.modules.tacotron2 import Tacotron, Decoder
    from multilingual_text_to_speech_b200.rng import MaskSource
    meta, sd, tape, inp, ref = _load(name)
    hp.reset()
    hp.load_state_dict(meta['hp'])
    model = Tacotron()
    model.load_state_dict(sd, strict=True)
    dev = torch.device('cuda:0')
    model = model.to(dev).eval()
    calls = []
    orig = F.decoder_forward_chunk
    F.decoder_forward_chunk = lambda *a, **k: (calls.append(a[-1]), orig(*a, **k))[1]
    old_chunk = Decoder.inference_chunk
    Decoder.inference_chunk = chunk
    MaskSource.use_tape(tape)
    try:
        language = inp['language'].to(dev) if 'language' in inp else None
        out = model.inference(inp['text'].to(dev), speaker=None, language=language)
    finally:
        MaskSource.use_tape(None)
        Decoder.inference_chunk = old_chunk
        F.decoder_forward_chunk = orig
    T = meta['T']
    assert tuple(out.shape) == tuple(ref.shape) == (hp.num_mels, T), (out.shape, ref.shape)
    assert_close(out, ref, 1e-3, 1e-4, f'{name}: inference spectrogram')
    # early exit: no more chunks were decoded than the stop rule needed (the reference stops after frame T of max_output_length)
    assert T < hp.max_output_length
    assert sum(calls) == min(-(-T // chunk) * chunk, hp.max_output_length), (calls, T)
