"""Input pipeline for the hot path: collate into the 8-tuple train.py consumes, and a pinned-memory prefetcher that overlaps the
host -> device copy of batch i+1 with the training step of batch i.

Reference: dataset/dataset.py:262-322 (`TextToSpeechCollate`).  Its sort branch cannot run as shipped (`languages.size(1)`
on a 1-D tensor and the `one_hots` / `one_hot` name mix-up at :294-303); the behaviour it was written for is implemented here:
sorting permutes every field consistently and the language ids stay ids (the model takes `LongTensor[B]`, modules/tacotron2.py:359-360).
Items are the dataset's 5-tuples `(speaker, language, utterance ids, mel [num_mels, frames], linear | None)`.
"""
import threading
import queue

import numpy as np
import torch


class TextToSpeechCollate:
    def __init__(self, sort_by_text_length, num_mels, stop_frames, multi_speaker, multi_language, pad_text_to=None, pad_frames_to=None):
        """pad_text_to / pad_frames_to: optional fixed paddings (bucketed batches keep one shape per bucket, so the library's workspaces
        and the CUDA-graph-friendly launch sequence are reused from step to step)."""
        self.sort = bool(sort_by_text_length)
        self.num_mels, self.stop_frames = int(num_mels), int(stop_frames)
        self.multi_speaker, self.multi_language = bool(multi_speaker), bool(multi_language)
        self.pad_text_to, self.pad_frames_to = pad_text_to, pad_frames_to

    def __call__(self, batch):
        n = len(batch)
        text_len = torch.tensor([len(item[2]) for item in batch], dtype=torch.long)
        mel_len = torch.tensor([np.shape(item[3])[1] for item in batch], dtype=torch.long)
        order = torch.argsort(text_len, descending=True, stable=True) if self.sort else torch.arange(n)
        L = int(self.pad_text_to or text_len.max())
        T = int(self.pad_frames_to or mel_len.max())
        if L < int(text_len.max()) or T < int(mel_len.max()):
            raise ValueError(f'fixed padding ({L}, {T}) is shorter than the longest item ({int(text_len.max())}, {int(mel_len.max())})')
        text = torch.zeros(n, L, dtype=torch.long)
        mel = torch.zeros(n, self.num_mels, T, dtype=torch.float32)
        stop = torch.zeros(n, T, dtype=torch.float32)
        for row, src in enumerate(order.tolist()):
            _, _, ids, m, _ = batch[src]
            text[row, :len(ids)] = torch.as_tensor(ids, dtype=torch.long)
            frames = np.shape(m)[1]
            mel[row, :, :frames] = torch.as_tensor(np.asarray(m), dtype=torch.float32)
            stop[row, max(frames - self.stop_frames, 0):] = 1.0          # dataset.py:320: ones from the last stop_frames real frames on
        speakers = torch.tensor([batch[i][0] for i in order.tolist()], dtype=torch.long) if self.multi_speaker else None
        languages = torch.tensor([batch[i][1] for i in order.tolist()], dtype=torch.long) if self.multi_language else None
        return text, text_len[order], mel, None, mel_len[order], stop, speakers, languages


class PinnedPrefetcher:
    """Iterates `loader` (any iterable of tuples of CPU tensors / None) and yields the same tuples ON THE DEVICE.

    A worker thread stages every batch in page-locked buffers; the copy of batch i+1 is enqueued on a side CUDA stream while the caller
    trains on batch i, and `__next__` only makes the compute stream wait on that copy's event (no host synchronisation).  `depth`
    batches are in flight.  This is the path bench.py's `e2e` number measures one step of (pinned H2D + step + D2H of the loss)."""

    def __init__(self, loader, device, depth=2):
        self.loader, self.device, self.depth = loader, torch.device(device), max(1, int(depth))
        if self.device.type != 'cuda':
            raise RuntimeError('PinnedPrefetcher feeds a CUDA device (the hot path has no CPU fallback)')

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        stream = torch.cuda.Stream(device=self.device)
        stop = threading.Event()

        def stage():
            try:
                for batch in self.loader:
                    if stop.is_set():
                        return
                    pinned = tuple(None if t is None else (t if t.is_pinned() else t.contiguous().pin_memory()) for t in batch)
                    with torch.cuda.stream(stream):
                        dev = tuple(None if t is None else t.to(self.device, non_blocking=True) for t in pinned)
                        ready = torch.cuda.Event()
                        ready.record(stream)
                    q.put((dev, pinned, ready))          # `pinned` is kept alive until the copy has been consumed
                q.put(None)
            except BaseException as exc:      # noqa: BLE001 -- surfaced in the consumer thread
                q.put(exc)

        worker = threading.Thread(target=stage, daemon=True)
        worker.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                dev, pinned, ready = item
                torch.cuda.current_stream(self.device).wait_event(ready)
                for t in dev:
                    if t is not None:
                        t.record_stream(torch.cuda.current_stream(self.device))
                yield dev
        finally:
            stop.set()
