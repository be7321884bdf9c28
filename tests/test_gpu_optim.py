"""Fused clip + Adam step (b200tts_adam_clip_step) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam + StepLR."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()


@pytest.mark.parametrize('max_norm', [None, 0.25])
def test_fused_adam_matches_torch(max_norm):
    from multilingual_text_to_speech_b200.optim import FlatParams, FusedAdam
    from multilingual_text_to_speech_b200.distributed import GradBucket
    dev = torch.device('cuda:0')
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.Tanh(), torch.nn.Linear(129, 5)).to(dev)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.Tanh(), torch.nn.Linear(129, 5)).to(dev)
    ref.load_state_dict(model.state_dict())
    flat, bucket = FlatParams(model), GradBucket(model, 1)
    opt = FusedAdam(flat, bucket, lr=1e-2, weight_decay=1e-3, max_grad_norm=max_norm, lr_decay_every=2, lr_decay=0.5)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-3)
    rsched = torch.optim.lr_scheduler.StepLR(ropt, 2, 0.5)
    g = torch.Generator().manual_seed(5)
    for it in range(5):
        x, y = torch.randn(16, 37, generator=g).to(dev), torch.randn(16, 5, generator=g).to(dev)
        bucket.zero()
        (model(x) - y).pow(2).mean().mul(30.0).backward()
        ropt.zero_grad()
        (ref(x) - y).pow(2).mean().mul(30.0).backward()
        rnorm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm) if max_norm else None
        info = opt.step()
        ropt.step(); rsched.step()
        if max_norm:
            assert abs(float(info[0]) - float(rnorm)) < 1e-4 * float(rnorm)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (it, float((p - q).abs().max()))
    assert model[0].weight.data_ptr() >= flat.flat.data_ptr()          # parameters live inside the flat buffer
