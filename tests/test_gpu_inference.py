"""Tacotron.inference (synthesize.py entry point, reference tacotron2.py:387-408): free-running decode inside the fused op, trimmed
at the stop token like the reference's loop.  Checked against the same model's eval-mode forward with a zero teacher-forcing ratio
(whose free-running path is pinned to the reference by the `lj_eval_free` golden case)."""
import pytest
import torch

import model_cases
from helpers import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()


def test_inference_matches_free_running_forward():
    from multilingual_text_to_speech_b200.rng import MaskSource
    from multilingual_text_to_speech_b200.params.params import Params as hp
    from multilingual_text_to_speech_b200.modules.tacotron2 import Decoder
    g = Golden('lj_eval_free')
    dev = torch.device('cuda:0')
    model = model_cases.build_model(g, dev).eval()
    T = 24
    model._decoder._max_frames = T
    text = g.inputs['text'][0].to(dev)
    L = int(text.shape[0])
    MaskSource.manual_seed(77)
    post_inf = model.inference(text.clone())
    # the same decode through forward(): eval mode, teacher forcing ratio 0 -> every frame free-running; same mask stream
    MaskSource.manual_seed(77)
    with torch.no_grad():
        post, pre, stop, align, _, _ = model(text[None], torch.tensor([L], device=dev), torch.zeros(1, hp.num_mels, T, device=dev),
                                             torch.tensor([T], device=dev), None, None, 0.0)
    cut = Decoder._stop_cut(stop[0].float().cpu(), hp.stop_frames)
    assert post_inf.shape == (hp.num_mels, cut), (post_inf.shape, cut)
    if cut == T:       # nothing trimmed: the post-net saw the same frames
        assert torch.allclose(post_inf, post[0], rtol=1e-4, atol=1e-5)
    assert torch.isfinite(post_inf).all()
