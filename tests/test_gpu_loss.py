"""Fused TacotronLoss op (csrc/loss.cu, forward + backward) against the CPU oracle's restatement of
modules/tacotron2.py:439-485 in fp64: loss terms, and the gradients w.r.t. pre / post / stop / alignment."""
import types
import pytest
import torch

from helpers import assert_close
from oracle import tacotron_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()


@pytest.mark.parametrize('B,N,T,L,g,guided', [(3, 12, 17, 11, 0.2, True), (10, 80, 301, 77, 0.2, True), (4, 80, 64, 180, 0.05, True),
                                             (2, 5, 9, 4, 0.2, False)])
def test_fused_loss_matches_oracle(B, N, T, L, g, guided):
    from multilingual_text_to_speech_b200 import functional as F
    gen = torch.Generator().manual_seed(B * 100 + T)
    pre, post, tgt = (torch.randn(B, N, T, generator=gen) for _ in range(3))
    stop = torch.randn(B, T, generator=gen) * 3
    tlen = torch.randint(max(1, L // 2), L + 1, (B,), generator=gen); tlen[0] = L
    mlen = torch.randint(max(1, T // 2), T + 1, (B,), generator=gen); mlen[0] = T
    for b in range(B):
        stop[b, mlen[b]:] = 1000.0                    # Tacotron.forward fills the padded stop logits (tacotron2.py:380)
    stop_t = torch.zeros(B, T)
    for b in range(B):
        stop_t[b, max(int(mlen[b]) - 3, 0):] = 1.0
    align = torch.softmax(torch.randn(B, T, L, generator=gen), dim=2)
    hp = types.SimpleNamespace(num_mels=N, reversal_classifier=False, guided_attention_loss=True)
    ref_in = [t.double().clone().requires_grad_(True) for t in (pre, post, stop, align)]
    total_o, parts_o = O.tacotron_loss(hp, g, tlen, mlen, ref_in[0], tgt.double(), ref_in[1], tgt.double(), ref_in[2], stop_t.double(), ref_in[3],
                                       guided=guided)
    wts = torch.tensor([0.7, 1.3, 2.0, 0.5], dtype=torch.float64)       # unequal upstream gradients exercise grad_losses
    names = ['mel_pre', 'mel_pos', 'stop_token', 'guided_att']
    sum(wts[k] * parts_o[n] for k, n in enumerate(names) if n in parts_o).backward()
    dev = torch.device('cuda:0')
    cu_in = [t.to(dev).clone().requires_grad_(True) for t in (pre, post, stop, align)]
    terms = F.tacotron_loss(cu_in[0], cu_in[1], cu_in[2], cu_in[3] if guided else None, tgt.to(dev), tgt.to(dev), stop_t.to(dev), tlen.to(dev),
                            mlen.to(dev), guided, g)
    for k, n in enumerate(names):
        ref = float(parts_o[n]) if n in parts_o else 0.0
        assert abs(float(terms[k]) - ref) < 1e-5 * max(1.0, abs(ref)), (n, float(terms[k]), ref)
    (terms * wts.float().to(dev)).sum().backward()
    for name, got, ref in zip(('pre', 'post', 'stop', 'align'), cu_in, ref_in):
        if name == 'align' and not guided:
            assert got.grad is None or float(got.grad.abs().max()) == 0.0
            continue
        scale = float(ref.grad.abs().max()) + 1e-30
        assert_close(got.grad, ref.grad, 1e-4, 1e-6 * scale, 'd_' + name)
