"""Packed bidirectional LSTM of the vanilla encoder (reference modules/encoder.py:33,41-44) against torch.nn.LSTM on a
pack_padded_sequence (the reference's own call, CPU fp64): outputs, input gradient and all eight parameter gradients.
Covers the persistent cluster kernel (bilstm_persist.cu: H % 32 == 0, H <= 256) at the real encoder dimensions and ragged lengths,
its per-step fallback (other H, or B200TTS_BILSTM_CHAIN=1), groups of utterances that end early, and B not a multiple of 8."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(x, lengths, params, dout):
    B, L, E = x.shape
    H = params[1].shape[1]
    lstm = torch.nn.LSTM(E, H, batch_first=True, bidirectional=True).double()
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0', 'weight_ih_l0_reverse', 'weight_hh_l0_reverse',
             'bias_ih_l0_reverse', 'bias_hh_l0_reverse']
    with torch.no_grad():
        for n, p in zip(names, params):
            getattr(lstm, n).copy_(p.double())
    xr = x.double().clone().requires_grad_(True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xr, lengths.cpu(), batch_first=True, enforce_sorted=False)
    y, _ = lstm(packed)
    y, _ = torch.nn.utils.rnn.pad_packed_sequence(y, batch_first=True, total_length=L)
    (y * dout.double()).sum().backward()
    return y.detach(), xr.grad, [getattr(lstm, n).grad for n in names]


def _run(B, L, E, H, lengths, seed=0):
    from multilingual_text_to_speech_b200 import functional as F
    from helpers import assert_close
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, E, generator=g)
    k = 1.0 / H ** 0.5
    shapes = [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)] * 2
    params = [(torch.rand(*s, generator=g) * 2 - 1) * k * 2.0 for s in shapes]      # twice PyTorch's init range: a livelier recurrence
    dout = torch.randn(B, L, 2 * H, generator=g)
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    y_ref, dx_ref, dp_ref = _reference(x, lengths, params, dout)
    dev = torch.device('cuda:0')
    xg = x.to(dev).requires_grad_(True)
    pg = [p.to(dev).requires_grad_(True) for p in params]
    y = F.bilstm(xg, lengths.to(dev), pg)
    (y * dout.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(y, y_ref, 1e-3, 1e-4, 'bilstm output')
    for b in range(B):                                   # exact zeros beyond the length (pad_packed_sequence semantics)
        if int(lengths[b]) < L:
            assert float(y[b, int(lengths[b]):].abs().max()) == 0.0
    scale = float(dx_ref.abs().max())
    assert_close(xg.grad, dx_ref, 2e-3, 2e-4 * scale, 'bilstm dx')
    for name, got, ref in zip(('w_ih', 'w_hh', 'b_ih', 'b_hh', 'w_ih_r', 'w_hh_r', 'b_ih_r', 'b_hh_r'), pg, dp_ref):
        assert_close(got.grad, ref, 2e-3, 2e-4 * float(ref.abs().max()), f'bilstm d{name}')


CASES = [
    # B, L, E, H, lengths
    (16, 40, 512, 256, None),                                            # cfg 1 encoder (LJ Speech defaults): cluster of 8
    (64, 60, 256, 128, None),                                            # cfg 3 encoder (shared_switching): cluster of 4
    (11, 23, 64, 32, [23, 23, 20, 17, 17, 9, 9, 9, 3, 1, 1]),            # cluster of 1, B % 8 != 0, a group that ends early
    (9, 19, 48, 24, [19, 12, 12, 11, 7, 5, 3, 2, 1]),                    # H % 32 != 0: per-step chain
]


@pytest.mark.parametrize('B,L,E,H,lengths', CASES)
def test_bilstm_matches_torch_packed_lstm(B, L, E, H, lengths):
    if lengths is None:
        g = torch.Generator().manual_seed(B * 131 + L)
        lengths = torch.randint(1, L + 1, (B,), generator=g).sort(descending=True).values.tolist()
        lengths[0] = L
    _run(B, L, E, H, lengths)


def test_bilstm_chain_switch_agrees():
    """A/B: the per-step chain (B200TTS_BILSTM_CHAIN=1) and the persistent kernel give the same result on the same inputs."""
    from multilingual_text_to_speech_b200 import functional as F
    g = torch.Generator().manual_seed(5)
    B, L, E, H = 12, 30, 128, 64
    dev = torch.device('cuda:0')
    x = torch.randn(B, L, E, generator=g).to(dev)
    params = [((torch.rand(*s, generator=g) * 2 - 1) * 0.2).to(dev) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)] * 2]
    lengths = torch.tensor([30, 30, 28, 25, 25, 20, 14, 14, 9, 5, 2, 1], device=dev)
    outs = []
    for chain in (False, True):
        if chain:
            os.environ['B200TTS_BILSTM_CHAIN'] = '1'
        try:
            xs = x.clone().requires_grad_(True)
            ps = [p.clone().requires_grad_(True) for p in params]
            y = F.bilstm(xs, lengths, ps)
            y.square().sum().backward()
            torch.cuda.synchronize()
            outs.append([y.detach(), xs.grad] + [p.grad for p in ps])
        finally:
            os.environ.pop('B200TTS_BILSTM_CHAIN', None)
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()) + 1e-7)
