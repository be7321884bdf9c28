"""Batch-1 synthesis latency of Tacotron.inference (chunked free-running decode with carried state and early exit, reference
modules/tacotron2.py:201-219, synthesize.py:81) on random-initialised weights: the stop token never fires, so exactly --frames frames
are decoded; reports ms per utterance and microseconds per mel frame (a 12.5 ms hop is 80 frames per second of audio)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='ljspeech')
    ap.add_argument('--text-len', type=int, default=120)
    ap.add_argument('--frames', type=int, default=640)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--iters', type=int, default=3)
    a = ap.parse_args()
    from multilingual_text_to_speech_b200 import configs, _lib
    from multilingual_text_to_speech_b200.modules.tacotron2 import Tacotron
    hp = configs.apply(a.config, max_output_length=a.frames)
    _lib.set_precision(a.precision)
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = Tacotron().to(dev).eval()
    with torch.no_grad():
        model._decoder._stop_prediction.bias.fill_(-100.0)      # random weights: keep the stop token from firing, decode all --frames frames
    text = torch.randint(1, hp.symbols_count() + 3, (a.text_len,), device=dev)
    language = None
    if hp.multi_language:
        language = torch.zeros(1, a.text_len, hp.language_number, device=dev)      # per-character language shares (synthesize.py one-hot form)
        language[:, :, 0] = 1.0
    speaker = torch.zeros(1, dtype=torch.long, device=dev) if hp.multi_speaker else None
    for it in range(a.iters + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            out = model.inference(text, speaker=speaker, language=language)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        T = out.shape[-1]
        print(f'iter {it}: {T} frames in {1e3 * dt:.1f} ms = {1e6 * dt / T:.1f} us / frame = {T / 80.0 / dt:.1f} x real time (12.5 ms hop)', flush=True)


if __name__ == '__main__':
    main()
