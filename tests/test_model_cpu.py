"""Host-side (CPU) checks of the module surface: constructor / state_dict compatibility with the reference."""
import pytest
import torch

from helpers import Golden, GOLDEN_CASES
import model_cases


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_state_dict_names_and_shapes_match_reference(name):
    g = Golden(name)
    model = model_cases.build_model(g)            # strict load: raises on any missing / unexpected key or shape mismatch
    own = model.state_dict()
    assert list(own.keys()) == list(g.sd.keys())
    for k, v in own.items():
        assert tuple(v.shape) == tuple(g.sd[k].shape), k
    # shared modules are registered twice, exactly like the reference
    assert model._decoder._prenet is model._prenet and model._decoder._attention is model._attention
    # the handles train.py reads (train.py:67,130,262-266)
    for attr in ('_encoder', '_decoder', '_postnet', '_prenet', '_embedding', '_attention'):
        assert len(list(getattr(model, attr).parameters())) > 0


def test_params_surface():
    from multilingual_text_to_speech_b200.params.params import Params as hp
    hp.reset()
    assert hp.symbols_count() == 70 and hp.decoder_dimension == 1024 and hp.attention_kernel_size == 31
    sd = hp.state_dict()
    assert 'batch_size' in sd and 'symbols_count' not in sd
    hp.load_state_dict({'encoder_type': 'generated'})
    assert hp.encoder_type == 'generated'
    hp.reset()
    assert hp.encoder_type == 'simple'


def test_no_cpu_fallback():
    from multilingual_text_to_speech_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('checks the no-GPU failure mode')
    g = Golden('lj_dropout')
    model = model_cases.build_model(g)
    i = g.inputs
    with pytest.raises(_lib.B200TTSError):
        model(i['text'], i['text_length'], i['target'], i['target_length'], None, None, 1.0)


def test_inference_stop_rule_matches_reference_loop():
    """Decoder._stop_cut reproduces where the reference's inference loop returns (modules/tacotron2.py:201-207): the frame on which
    the stop token fires for the (stop_frames + 1)-th time -- not necessarily consecutively -- else all frames."""
    import torch
    from multilingual_text_to_speech_b200.modules.tacotron2 import Decoder

    def reference_loop(stop_logits, stop_frames_hp):
        stop_frames = -1
        for i in range(len(stop_logits)):
            if torch.sigmoid(stop_logits[i]).ge(0.5):
                if stop_frames == -1:
                    stop_frames = stop_frames_hp
                    continue
                stop_frames -= 1
                if stop_frames == 0:
                    return i + 1
        return len(stop_logits)

    g = torch.Generator().manual_seed(11)
    for trial in range(100):
        x = torch.randn(80, generator=g) * 2 + (torch.arange(80) - 40) * 0.15
        for sf in (1, 5):
            assert Decoder._stop_cut(x, sf) == reference_loop(x, sf)
    assert Decoder._stop_cut(torch.full((10,), -3.0), 5) == 10                      # never fires: all frames
    assert Decoder._stop_cut(torch.tensor([-1.0, 2.0, 2.0, -1.0, 2.0, 2.0]), 3) == 6   # fires at 1 (arms), 2, 4, 5 -> cut after frame 5


def test_streaming_stop_rule_equals_the_loop_rule():
    """Decoder._StopRule fed chunk by chunk (the chunked inference) cuts exactly where the one-shot rule does."""
    import torch
    from multilingual_text_to_speech_b200.modules.tacotron2 import Decoder
    g = torch.Generator().manual_seed(5)
    for trial in range(200):
        n = int(torch.randint(1, 90, (1,), generator=g))
        x = torch.randn(n, generator=g) * 2 + (torch.arange(n) - n / 2) * 0.1
        for sf in (1, 5):
            want = Decoder._stop_cut(x, sf)
            chunk = int(torch.randint(1, 17, (1,), generator=g))
            rule, cut = Decoder._StopRule(sf), None
            for i in range(0, n, chunk):
                cut = rule.feed(x[i:i + chunk])
                if cut is not None:
                    break
            assert (cut if cut is not None else n) == want, (trial, sf, chunk, cut, want)
