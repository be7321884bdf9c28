"""Small packed bi-LSTM forward + backward on the persistent cluster kernels (for compute-sanitizer runs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multilingual_text_to_speech_b200 import functional as F  # noqa: E402

g = torch.Generator().manual_seed(0)
B, L, E, H = 11, 9, 64, 64
dev = torch.device('cuda:0')
x = torch.randn(B, L, E, generator=g).to(dev).requires_grad_(True)
params = [((torch.rand(*s, generator=g) * 2 - 1) * 0.2).to(dev).requires_grad_(True) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)] * 2]
lengths = torch.tensor([9, 9, 8, 7, 7, 5, 4, 4, 2, 1, 1], device=dev)
y = F.bilstm(x, lengths, params)
y.square().sum().backward()
torch.cuda.synchronize()
print('ok', float(y.abs().sum()), float(x.grad.abs().sum()))
