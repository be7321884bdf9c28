"""Golden vectors of the reference's INFERENCE path (build container only; imports the unmodified reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_inference.py

`Tacotron.inference(text, speaker, language)` (modules/tacotron2.py:387-408 -> Decoder.inference :216-219 -> the early-exit loop
:201-207), batch 1, eval mode, with the always-on prenet dropout masks recorded in call order (2 per decoder step).  Cases:
  inf_lj              monolingual default configuration (small dimensions)
  inf_generated_mix   generated encoder, per-character language weights [1, L, G] -> the language-mixing branch modules/encoder.py:213-219
                      (two languages blended on a stretch of characters, code-switching elsewhere)
"""
import os
import sys
import json
import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import SMALL      # noqa: E402

CASES = {
    'inf_lj': (dict(max_output_length=60), 9, 0.0),
    'inf_generated_mix': (dict(encoder_type='generated', multi_language=True, languages=['a', 'b', 'c'], language_embedding_dimension=6,
                               generator_dim=5, generator_bottleneck_dim=3, max_output_length=60, decoder_regularization='zoneout'), 11, 0.0),
}


def main():
    sys.path.insert(0, REF)
    import torch
    import torch.nn.functional as F
    import utils  # noqa: F401
    from params.params import Params as hp
    from modules.tacotron2 import Tacotron

    defaults = dict(hp.state_dict())
    real_dropout = F.dropout
    record = []

    def taped_dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        keep = (torch.rand(tuple(input.shape)) >= p).to(input.dtype)
        record.append(keep.detach().clone())
        return input * keep * (1.0 / (1.0 - p))

    for name, (over, L, _) in CASES.items():
        hp.load_state_dict(defaults)
        hp.load_state_dict(SMALL)
        hp.load_state_dict(over)
        hp.language_number = len(hp.languages) if hp.multi_language else 0
        hp.speaker_number = 0
        G = max(hp.language_number, 1)
        torch.manual_seed(sum(map(ord, name)))
        model = Tacotron()
        with torch.no_grad():
            model._attention._energy.weight.mul_(6.0)
            model._attention._memory.weight.mul_(3.0)
            model._attention._query.weight.mul_(3.0)
            for prm in model._decoder._attention_lstm.parameters():
                prm.mul_(2.0)
            model._decoder._stop_prediction.weight.mul_(4.0)
        model.eval()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        text = torch.randint(1, hp.symbols_count() + 3, (L,))
        language = None
        if hp.multi_language:
            lw = torch.zeros(1, L, G)
            lw[0, :4, 0] = 1.0                      # language a
            lw[0, 4:8, 1] = 0.7; lw[0, 4:8, 2] = 0.3    # blend of b and c (weights sum to 1)
            lw[0, 8:, 2] = 2.0                      # language c with an un-normalised weight
            language = lw
        # pass 1 (stop token disabled by a huge negative bias): the stop logits of the free-running trajectory under the seeded masks;
        # the bias is then placed at their 75th percentile so that the token fires intermittently and the exit lands mid-sequence
        # (the masks of pass 2 are the same draws: same seed, same call order up to the exit)
        def run(bias):
            with torch.no_grad():
                model._decoder._stop_prediction.bias.fill_(bias)
            record.clear()
            torch.manual_seed(99)
            F.dropout = taped_dropout
            try:
                with torch.no_grad():
                    return model.inference(text.clone(), speaker=None, language=language)
            finally:
                F.dropout = real_dropout
        captured = {}
        hook = model._decoder._stop_prediction.register_forward_hook(lambda m, i, o: captured.setdefault('v', []).append(float(o)))
        run(-1000.0)
        hook.remove()
        logits = torch.tensor(captured['v']) + 1000.0
        if logits[:30].mean() > logits[30:].mean():      # make the token more likely late than early (as in a trained model)
            with torch.no_grad():
                model._decoder._stop_prediction.weight.neg_()
            logits = -logits
            sd0['_decoder._stop_prediction.weight'] = model._decoder._stop_prediction.weight.detach().clone()
        out = run(-float(torch.quantile(logits, 0.75)))
        sd0['_decoder._stop_prediction.bias'] = model._decoder._stop_prediction.bias.detach().clone()
        T = out.shape[1]
        assert len(record) == 2 * T, (len(record), T)
        P = hp.prenet_dimension
        m0 = torch.stack([record[2 * i] for i in range(T)]).reshape(T, 1, P)
        m1 = torch.stack([record[2 * i + 1] for i in range(T)]).reshape(T, 1, P)
        res = {'meta': np.frombuffer(json.dumps(dict(
            hp={k: v for k, v in hp.state_dict().items() if isinstance(v, (int, float, str, bool, list))}, L=L, T=T)).encode(), dtype=np.uint8),
            'in.text': text.numpy(), 'tape.step_prenet0': m0.numpy().astype(np.uint8), 'tape.step_prenet1': m1.numpy().astype(np.uint8),
            'out.post': out.detach().numpy()}
        if language is not None:
            res['in.language'] = language.numpy()
        for k, v in sd0.items():
            res['sd.' + k] = v.numpy()
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **res)
        print(f'{name}: T={T} of max {hp.max_output_length}, {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
