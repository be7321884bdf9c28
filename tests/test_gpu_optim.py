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


def test_flat_layout_is_aligned_and_survives_set_to_none():
    from multilingual_text_to_speech_b200.distributed import GradBucket, flat_layout
    dev = torch.device('cuda:0')
    model = torch.nn.Sequential(torch.nn.Linear(3, 1), torch.nn.Linear(1, 7)).to(dev)      # 1-element bias in the middle of the buffer
    bucket = GradBucket(model, 1)
    offs, total = flat_layout(bucket.params)
    assert all(o % 4 == 0 for o in offs) and total % 4 == 0
    assert all(p.grad.data_ptr() % 16 == 0 for p in bucket.params)
    x = torch.randn(5, 3, device=dev)
    model(x).sum().backward()
    want = [p.grad.clone() for p in bucket.params]
    model.zero_grad(set_to_none=True)                      # detaches every .grad from the bucket
    bucket.zero()                                          # re-binds
    model(x).sum().backward()
    for p, w in zip(bucket.params, want):
        lo = bucket.flat.data_ptr()
        assert lo <= p.grad.data_ptr() < lo + bucket.flat.numel() * 4
        assert torch.equal(p.grad, w)
    # a gradient produced while detached is folded back in, not lost
    model.zero_grad(set_to_none=True)
    model(x).sum().backward()
    bucket.allreduce()
    for p, w in zip(bucket.params, want):
        assert torch.allclose(p.grad, 2 * w)


def test_fused_adam_on_the_real_model():
    """FlatParams + GradBucket + FusedAdam on a (small-dimension) Tacotron against clip_grad_norm_ + torch.optim.Adam, 3 steps."""
    import model_cases
    from helpers import Golden
    from multilingual_text_to_speech_b200.optim import FlatParams, FusedAdam
    from multilingual_text_to_speech_b200.distributed import GradBucket
    from multilingual_text_to_speech_b200.modules.tacotron2 import TacotronLoss
    from multilingual_text_to_speech_b200.rng import MaskSource
    from multilingual_text_to_speech_b200.params.params import Params as hp
    g = Golden('generated_training')
    dev = torch.device('cuda:0')
    model, ref = model_cases.build_model(g, dev), model_cases.build_model(g, dev)
    flat, bucket = FlatParams(model), GradBucket(model, 1)
    opt = FusedAdam(flat, bucket, lr=1e-3, weight_decay=1e-6, max_grad_norm=0.25)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-6)
    i = {k: v.to(dev) for k, v in g.inputs.items()}

    def loss_of(m):
        crit = TacotronLoss(hp.guided_attention_steps, g.meta['guided_g'], hp.guided_attention_gain)
        MaskSource.use_tape(g.tape)
        try:
            post, pre, stop, align, spk, enc = m(i['text'], i['text_length'], i['target'], i['target_length'], i.get('speakers'), i.get('languages'), 1.0)
        finally:
            MaskSource.use_tape(None)
        return crit(i['text_length'], i['target_length'], pre, i['target'], post, i['target'], stop, i['stop_target'], align, i.get('speakers'),
                    spk, enc, None)[0]

    for it in range(3):
        bucket.zero()
        loss_of(model).backward()
        ropt.zero_grad()
        loss_of(ref).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.25)
        opt.step(); ropt.step()
        worst = max(float((p - q).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
        assert worst < 2e-5, (it, worst)
    assert all(p.data_ptr() % 16 == 0 for p in model.parameters())


def test_pinned_prefetcher_delivers_batches_in_order():
    from multilingual_text_to_speech_b200.utils.data import PinnedPrefetcher
    batches = [(torch.full((4, 7), k, dtype=torch.long), None, torch.randn(4, 3, 9) + k) for k in range(6)]
    seen = []
    for dev in PinnedPrefetcher(batches, 'cuda:0', depth=2):
        assert dev[0].is_cuda and dev[1] is None and dev[2].is_cuda
        seen.append((int(dev[0][0, 0]), dev[2].cpu()))
    assert [k for k, _ in seen] == list(range(6))
    for (k, got), ref in zip(seen, batches):
        assert torch.equal(got, ref[2])


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (one CUDA graph per step) computes what the eager step computes: same loss and gradients on a mask tape, and
    fresh dropout masks on every replay without one (device-side mask epoch)."""
    import model_cases
    from helpers import Golden
    from multilingual_text_to_speech_b200.distributed import GradBucket
    from multilingual_text_to_speech_b200.graph import GraphedTrainStep
    from multilingual_text_to_speech_b200.modules.tacotron2 import TacotronLoss
    from multilingual_text_to_speech_b200.rng import MaskSource
    from multilingual_text_to_speech_b200.params.params import Params as hp
    g = Golden('generated_training')
    dev = torch.device('cuda:0')
    model = model_cases.build_model(g, dev)
    bucket = GradBucket(model, 1)
    crit = TacotronLoss(hp.guided_attention_steps, g.meta['guided_g'], hp.guided_attention_gain)
    batch = {k: v.to(dev) for k, v in g.inputs.items()}
    batch.setdefault('speakers', None); batch.setdefault('languages', None)
    # masks resident on the device (a host -> device copy from pageable memory cannot be captured); the teacher coins stay on the host
    MaskSource.use_tape({k: (v if k == 'teacher' else v.to(dev)) for k, v in g.tape.items()})
    step = None
    try:
        # (no autograd graph built on another stream may be alive at capture time: its AccumulateGrad nodes would run on that stream)
        step = GraphedTrainStep(model, crit, bucket, batch, teacher_forcing=1.0, warmup=2)
        got = []
        for _ in range(2):
            loss_g = step(batch)
            torch.cuda.synchronize()
            got.append((float(loss_g), bucket.flat.clone()))
        step.close(); step = None
        bucket.zero()
        post, pre, stop, align, spk, enc = model(batch['text'], batch['text_length'], batch['target'], batch['target_length'], batch['speakers'],
                                                 batch['languages'], 1.0)
        loss, _ = crit(batch['text_length'], batch['target_length'], pre, batch['target'], post, batch['target'], stop, batch['stop_target'],
                       align, batch['speakers'], spk, enc, None)
        loss.backward()
        want_loss, want_grad = float(loss), bucket.flat.clone()
        del loss, post, pre, stop, align, spk, enc
        for got_loss, got_grad in got:
            assert abs(got_loss - want_loss) < 1e-6 * max(1.0, abs(want_loss))
            assert torch.allclose(got_grad, want_grad, rtol=1e-5, atol=1e-8)
    finally:
        MaskSource.use_tape(None)
        if step is not None:
            step.close()
    # without a tape: every replay draws new masks (the loss changes from replay to replay)
    MaskSource.manual_seed(5)
    step = GraphedTrainStep(model, crit, bucket, batch, teacher_forcing=1.0, warmup=2)
    try:
        losses = []
        for _ in range(3):
            losses.append(float(step(batch)))
        assert len(set(losses)) == 3, losses
    finally:
        step.close()
