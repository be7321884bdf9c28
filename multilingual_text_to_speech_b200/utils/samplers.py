"""Batch samplers of the input pipeline (reference utils/samplers.py:6-122) plus a length-bucketed variant.

The grouped encoders need every mini-batch laid out so that position ``i + k * L`` holds a sample of language ``i``
(``L`` languages), which lets ``[B, C, T]`` be viewed as ``[B / L, L * C, T]`` with one convolution group per language
(reference samplers.py:50-66).  ``PerfectBatchSampler`` reproduces the reference sampler;
``BucketedPerfectBatchSampler`` keeps that layout and additionally draws the utterances of one batch from the same text-length
bucket, which cuts the padding of the fused decoder's [B, T] work.  With one process per GPU every rank
takes a contiguous slice of the global batch (``shard``): the slice keeps the language layout because its size is a multiple of L.

``data_source`` only needs ``len()`` and ``data_source.items[idx]['language']`` (an int in ``range(len(languages))``).
"""
import random

import torch
from torch.utils.data.sampler import Sampler, WeightedRandomSampler, SubsetRandomSampler


def _language_indices(data_source):
    label_indices = {}
    for idx in range(len(data_source)):
        label_indices.setdefault(data_source.items[idx]['language'], []).append(idx)
    return label_indices


class RandomImbalancedSampler(Sampler):
    """Samples an imbalanced dataset randomly with repetition, every language equally likely (samplers.py:6-30)."""

    def __init__(self, data_source):
        freq = {k: len(v) for k, v in _language_indices(data_source).items()}
        total = float(sum(freq.values()))
        weights = [total / freq[data_source.items[idx]['language']] for idx in range(len(data_source))]
        self._sampler = WeightedRandomSampler(weights, len(weights))

    def __iter__(self):
        return iter(self._sampler)

    def __len__(self):
        return len(self._sampler)


class SubsetSampler(Sampler):
    """Samples elements sequentially from a given list of indices (samplers.py:33-47)."""

    def __init__(self, indices):
        self.indices = indices

    def __iter__(self):
        return (self.indices[i] for i in range(len(self.indices)))

    def __len__(self):
        return len(self.indices)


def _trim_incomplete(batch, n_languages, dp_devices):
    """Last, incomplete batch: keep a whole number of language groups that is divisible by the device count (samplers.py:100-108)."""
    groups = len(batch) // n_languages
    groups = (groups // dp_devices) * dp_devices
    return batch[:groups * n_languages]


class PerfectBatchSampler(Sampler):
    """Mini-batches for the grouped encoders: language ``i`` sits at positions ``i + k * L`` (samplers.py:50-122)."""

    def __init__(self, data_source, languages, batch_size, data_parallel_devices=1, shuffle=True, drop_last=False):
        assert batch_size % (len(languages) * data_parallel_devices) == 0, \
            'Batch size must be divisible by number of languages times the number of data parallel devices (if enabled).'
        label_indices = _language_indices(data_source)
        make = SubsetRandomSampler if shuffle else SubsetSampler
        self._samplers = [make(label_indices.get(i, [])) for i, _ in enumerate(languages)]
        self._batch_size = batch_size
        self._drop_last = drop_last
        self._dp_devices = data_parallel_devices

    def __iter__(self):
        batch = []
        iters = [iter(s) for s in self._samplers]
        while True:
            group = []
            for it in iters:
                idx = next(it, None)
                if idx is None:
                    break
                group.append(idx)
            if len(group) < len(iters):
                break
            batch += group
            if len(batch) == self._batch_size:
                yield batch
                batch = []
        if not self._drop_last and batch:
            batch = _trim_incomplete(batch, len(self._samplers), self._dp_devices)
            if batch:
                yield batch

    def __len__(self):
        per_language = self._batch_size // len(self._samplers)
        return min((len(s) + per_language - 1) // per_language for s in self._samplers)


class BucketedPerfectBatchSampler(Sampler):
    """Language-balanced mini-batches (layout of ``PerfectBatchSampler``) whose utterances have similar text lengths.

    Per language the (shuffled) indices are cut into buckets of ``bucket_batches`` batches' worth of samples, each bucket is sorted by
    length and cut into per-batch groups; the groups of every language are then ordered by mean length and groups of equal rank are
    interleaved into one batch, so that all languages of a batch come from the same length range.  Batch order is shuffled.

    Arguments:
        data_source -- dataset with ``items[idx]['language']``
        languages -- list of languages to sample from (index = language id)
        batch_size -- global batch size, divisible by ``len(languages) * data_parallel_devices``
        lengths -- sequence or callable idx -> text length
        bucket_batches -- batches per length bucket (1 = no sorting window, larger = tighter lengths, less randomness)
        data_parallel_devices -- number of ranks the global batch is sharded over (see ``shard``)
        shuffle, drop_last, seed -- as usual; the epoch is mixed into the seed by ``set_epoch``
    """

    def __init__(self, data_source, languages, batch_size, lengths, bucket_batches=8, data_parallel_devices=1, shuffle=True,
                 drop_last=False, seed=0):
        self._L = len(languages)
        assert batch_size % (self._L * data_parallel_devices) == 0, \
            'Batch size must be divisible by number of languages times the number of data parallel devices (if enabled).'
        self._label_indices = _language_indices(data_source)
        self._length = lengths if callable(lengths) else (lambda i, _l=lengths: _l[i])
        self._batch_size, self._per_language = batch_size, batch_size // self._L
        self._bucket = max(1, int(bucket_batches)) * self._per_language
        self._dp_devices, self._shuffle, self._drop_last = data_parallel_devices, shuffle, drop_last
        self._seed, self._epoch = seed, 0

    def set_epoch(self, epoch):
        self._epoch = int(epoch)

    def _groups_of_language(self, lang, rng):
        idx = list(self._label_indices.get(lang, []))
        if self._shuffle:
            rng.shuffle(idx)
        groups = []
        for s in range(0, len(idx), self._bucket):
            bucket = sorted(idx[s:s + self._bucket], key=self._length)
            groups += [bucket[g:g + self._per_language] for g in range(0, len(bucket), self._per_language)]
        return groups

    def __iter__(self):
        rng = random.Random(self._seed * 1000003 + self._epoch)
        per_lang = [self._groups_of_language(lang, rng) for lang in range(self._L)]
        n_full = min(sum(len(g) == self._per_language for g in groups) for groups in per_lang)
        batches = []
        ranked = []
        for groups in per_lang:
            full = [g for g in groups if len(g) == self._per_language]
            full.sort(key=lambda g: sum(self._length(i) for i in g) / len(g))
            # drop the surplus groups of the larger languages evenly over the length range
            keep = [full[(j * len(full)) // n_full] for j in range(n_full)] if n_full else []
            ranked.append(keep)
        for j in range(n_full):
            batch = []
            for k in range(self._per_language):
                batch += [ranked[lang][j][k] for lang in range(self._L)]          # position lang + k * L
            batches.append(batch)
        if not self._drop_last:
            # one last, smaller batch from the leftovers of every language (as many whole language groups as all languages can fill)
            used = [set(i for g in ranked[lang] for i in g) for lang in range(self._L)]
            rest = [[i for i in self._label_indices.get(lang, []) if i not in used[lang]] for lang in range(self._L)]
            k_max = min(len(r) for r in rest) if rest else 0
            k_max = min(k_max, self._per_language - 1)
            tail = []
            for k in range(k_max):
                tail += [sorted(rest[lang], key=self._length)[k] for lang in range(self._L)]
            tail = _trim_incomplete(tail, self._L, self._dp_devices)
            if tail:
                batches.append(tail)
        if self._shuffle:
            rng.shuffle(batches)
        return iter(batches)

    def __len__(self):
        return sum(1 for _ in iter(self))


def shard(batch, rank, world_size, n_languages):
    """Contiguous slice of a global batch for one rank; keeps the ``i + k * L`` language layout."""
    assert len(batch) % (world_size * n_languages) == 0, 'global batch is not divisible by ranks x languages'
    per_rank = len(batch) // world_size
    return batch[rank * per_rank:(rank + 1) * per_rank]
