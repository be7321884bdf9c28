"""Batch samplers (reference utils/samplers.py:50-122 semantics + the length-bucketed variant): language layout, divisibility,
last-batch handling, length tightness, per-rank sharding."""
import random
import types

import pytest

from multilingual_text_to_speech_b200.utils.samplers import (PerfectBatchSampler, BucketedPerfectBatchSampler, SubsetSampler,
                                                             RandomImbalancedSampler, shard)


def _dataset(counts, seed=0):
    rng = random.Random(seed)
    items = [{'language': lang, 'len': rng.randint(10, 200)} for lang, n in enumerate(counts) for _ in range(n)]
    rng.shuffle(items)
    return types.SimpleNamespace(items=items, __len__=None), items


class _DS:
    def __init__(self, items):
        self.items = items

    def __len__(self):
        return len(self.items)


def _check_layout(batch, items, L):
    assert len(batch) % L == 0
    for pos, idx in enumerate(batch):
        assert items[idx]['language'] == pos % L, (pos, idx)


@pytest.mark.parametrize('shuffle', [False, True])
def test_perfect_batch_sampler_layout_and_tail(shuffle):
    _, items = _dataset([13, 20, 17])
    ds = _DS(items)
    s = PerfectBatchSampler(ds, [0, 1, 2], 12, data_parallel_devices=2, shuffle=shuffle, drop_last=False)
    batches = list(s)
    # 13 samples of the rarest language -> 3 full batches (4 per language) and a tail of 1 group, trimmed to 0 (not divisible by 2)
    assert [len(b) for b in batches] == [12, 12, 12]
    for b in batches:
        _check_layout(b, items, 3)
    seen = [i for b in batches for i in b]
    assert len(seen) == len(set(seen))
    assert len(s) == 4                                   # the reference's __len__ counts the (possibly dropped) tail too
    s1 = PerfectBatchSampler(ds, [0, 1, 2], 12, data_parallel_devices=1, shuffle=shuffle, drop_last=False)
    assert [len(b) for b in s1] == [12, 12, 12, 3]
    s2 = PerfectBatchSampler(ds, [0, 1, 2], 12, shuffle=shuffle, drop_last=True)
    assert [len(b) for b in s2] == [12, 12, 12]


def test_perfect_batch_sampler_sequential_order():
    items = [{'language': i % 2} for i in range(10)]
    s = PerfectBatchSampler(_DS(items), [0, 1], 4, shuffle=False)
    assert list(s) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    with pytest.raises(AssertionError):
        PerfectBatchSampler(_DS(items), [0, 1], 6, data_parallel_devices=2)


def test_bucketed_sampler_layout_lengths_and_determinism():
    _, items = _dataset([400, 520, 460, 410, 433], seed=3)
    ds = _DS(items)
    L, B = 5, 40
    lengths = [it['len'] for it in items]
    s = BucketedPerfectBatchSampler(ds, list(range(L)), B, lengths, bucket_batches=8, data_parallel_devices=4, seed=7)
    batches = list(s)
    assert len(batches) == len(s) and len(batches) >= 400 // (B // L)
    seen = [i for b in batches for i in b]
    assert len(seen) == len(set(seen))
    spreads = []
    for b in batches:
        _check_layout(b, items, L)
        assert len(b) % (L * 4) == 0
        spreads.append(max(lengths[i] for i in b) - min(lengths[i] for i in b))
    plain = list(PerfectBatchSampler(ds, list(range(L)), B, shuffle=True))
    plain_spread = sum(max(lengths[i] for i in b) - min(lengths[i] for i in b) for b in plain) / len(plain)
    assert sum(spreads) / len(spreads) < 0.5 * plain_spread            # bucketing halves the in-batch length spread (at least)
    # same seed and epoch -> same batches; another epoch -> different order
    again = list(BucketedPerfectBatchSampler(ds, list(range(L)), B, lengths, bucket_batches=8, data_parallel_devices=4, seed=7))
    assert again == batches
    s.set_epoch(1)
    assert list(s) != batches


def test_shard_keeps_language_layout():
    _, items = _dataset([64, 64, 64, 64])
    ds = _DS(items)
    batch = next(iter(PerfectBatchSampler(ds, [0, 1, 2, 3], 32, data_parallel_devices=4)))
    parts = [shard(batch, r, 4, 4) for r in range(4)]
    assert sum(parts, []) == batch
    for p in parts:
        _check_layout(p, items, 4)


def test_small_samplers():
    assert list(SubsetSampler([5, 3, 9])) == [5, 3, 9]
    items = [{'language': 0}] * 90 + [{'language': 1}] * 10
    r = RandomImbalancedSampler(_DS(items))
    draws = list(r)
    assert len(draws) == 100 and 25 < sum(items[i]['language'] for i in draws) < 75       # both languages about equally likely


def test_collate_matches_the_reference_layout_and_fixes_the_sort_branch():
    """TextToSpeechCollate (dataset/dataset.py:262-322): padding, stop targets from the last `stop_frames` real frames on, and the
    sort branch (broken in the reference) permuting every field consistently."""
    import numpy as np
    from multilingual_text_to_speech_b200.utils.data import TextToSpeechCollate
    rng = np.random.default_rng(0)
    items = []
    for k, (n, f) in enumerate([(5, 30), (9, 44), (3, 12), (9, 40)]):
        items.append((k % 3, k % 2, list(rng.integers(1, 50, n)), rng.standard_normal((8, f)).astype(np.float32), None))
    plain = TextToSpeechCollate(False, 8, 5, True, True)(items)
    text, tl, mel, lin, ml, stop, spk, lang = plain
    assert text.shape == (4, 9) and mel.shape == (4, 8, 44) and lin is None and stop.shape == (4, 44)
    assert tl.tolist() == [5, 9, 3, 9] and ml.tolist() == [30, 44, 12, 40] and spk.tolist() == [0, 1, 2, 0] and lang.tolist() == [0, 1, 0, 1]
    assert text[0, 5:].eq(0).all() and (mel[2, :, 12:] == 0).all() and np.allclose(mel[2, :, :12].numpy(), items[2][3])
    assert stop[0, :25].eq(0).all() and stop[0, 25:].eq(1).all()            # ones from frame 30 - 5 on, including the padding
    srt = TextToSpeechCollate(True, 8, 5, True, True)(items)
    assert srt[1].tolist() == [9, 9, 5, 3] and srt[4].tolist() == [44, 40, 30, 12]      # stable: item 1 before item 3
    assert srt[6].tolist() == [1, 0, 0, 2] and srt[7].tolist() == [1, 1, 0, 0]
    assert np.allclose(srt[2][3, :, :12].numpy(), items[2][3]) and srt[0][2, :5].tolist() == [int(v) for v in items[0][2]]
    fixed = TextToSpeechCollate(False, 8, 5, False, False, pad_text_to=16, pad_frames_to=64)(items)
    assert fixed[0].shape == (4, 16) and fixed[2].shape == (4, 8, 64) and fixed[6] is None and fixed[7] is None
