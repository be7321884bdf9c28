"""Dropout-mask source for the host modules.

Normal operation: masks come from the library's counter-based generator (one stream id per draw).
Parity tests: a *mask tape* (dict key -> uint8/float tensor) replays the exact masks the reference drew
(SURVEY.md D8 / appendix A.7); a site whose key is absent from an active tape gets no dropout.
"""
import torch


class MaskSource:
    seed = 1234
    counter = 0
    tape = None

    @classmethod
    def manual_seed(cls, seed):
        cls.seed, cls.counter = int(seed), 0

    @classmethod
    def use_tape(cls, tape):
        cls.tape = tape

    @classmethod
    def keep_mask(cls, key, shape, rate, device):
        if cls.tape is not None:
            t = cls.tape.get(key)
            if t is None:
                return None
            t = t.to(device=device, dtype=torch.uint8).contiguous()
            assert tuple(t.shape) == tuple(shape), (key, tuple(t.shape), tuple(shape))
            return t
        if rate <= 0.0:
            return None
        from . import functional as F
        cls.counter += 1
        return F.fill_keep_mask(tuple(shape), rate, cls.seed, cls.counter, device)

    @classmethod
    def raw(cls, key):
        return None if cls.tape is None else cls.tape.get(key)
