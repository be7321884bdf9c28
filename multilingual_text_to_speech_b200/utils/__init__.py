"""Helpers with the reference's names (reference utils/__init__.py:7-37)."""
from collections import OrderedDict

import torch

from ..params.params import Params as hp


def lengths_to_mask(lengths, max_length=None):
    """Boolean mask [B, max_length] with True at positions < length."""
    ml = torch.max(lengths) if max_length is None else max_length
    return torch.arange(ml, device=lengths.device)[None, :] < lengths[:, None]


def to_gpu(x):
    if x is None:
        return x
    x = x.contiguous()
    return x.cuda(non_blocking=True) if torch.cuda.is_available() else x


def remove_dataparallel_prefix(state_dict):
    return OrderedDict((k[7:] if k.startswith('module.') else k, v) for k, v in state_dict.items())


def build_model(checkpoint, force_cpu=False):
    """Load hyper-parameters and weights from a reference-format checkpoint and build the model."""
    from ..modules.tacotron2 import Tacotron
    if force_cpu or not torch.cuda.is_available():
        raise RuntimeError('the B200-native Tacotron has no CPU path; a CUDA device is required')
    state = torch.load(checkpoint, map_location='cuda')
    hp.load_state_dict(state['parameters'])
    model = Tacotron()
    model.load_state_dict(remove_dataparallel_prefix(state['model']))
    return model.to('cuda')
