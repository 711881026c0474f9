"""Batched long-form inference up to token ids (BASELINE configs[3]).

The reference transcribes one file at a time without a padding mask (inference.py:101-148: preprocessor ->
model(features) -> argmax -> tokenizer.ctc_decode).  Here a whole batch of waveforms goes through the GPU
front-end, the encoder with the key-padding mask of the training path, and the argmax/collapse kernel."""
from typing import List, Optional

import torch

from . import _lib as L
from .data.preprocessing import AudioPreprocessor
from .utils.decoding import GreedyDecoder


class _Job:
    """Result handle of BatchedInference.submit()."""

    def __init__(self, tokens, lengths, done, pool):
        self._tokens, self._lengths, self._done, self._pool = tokens, lengths, done, pool

    def result(self) -> List[List[int]]:
        self._done.synchronize()
        out = [self._tokens[b, : int(self._lengths[b])].tolist() for b in range(self._tokens.shape[0])]
        if self._pool is not None:  # hand the pinned buffers back for the next batch
            self._pool.append((self._tokens, self._lengths))
            self._pool = None
        return out


class BatchedInference:
    def __init__(self, model, preprocessor: Optional[AudioPreprocessor] = None, tokenizer=None, blank_id: int = 0):
        self.model = model.eval()
        self.preprocessor = preprocessor if preprocessor is not None else AudioPreprocessor(device="cuda")
        self.decoder = GreedyDecoder(tokenizer, blank_id=blank_id)
        self._copy_stream = None
        self._staged = []  # device buffers of the batches staged ahead (two are enough for a prefetch depth of one)
        self._host_free = []  # pinned (tokens, lengths) buffers returned by finished jobs

    def stage(self, waves_host: torch.Tensor):
        """Start the host -> device copy of a (pinned) waveform batch on a copy stream and return a handle for
        logits() / transcribe_ids().  Staging batch i+1 before consuming batch i overlaps its transfer (123 MB for
        32 x 60 s) with the encoder of batch i."""
        dev = next(self.model.parameters()).device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        buf = None
        for cand in self._staged:  # reuse a buffer whose previous consumer has been enqueued and finished
            if cand["free"].query() and cand["buf"].numel() >= waves_host.numel():
                buf = cand
                break
        if buf is None:
            buf = {"buf": torch.empty(waves_host.numel(), dtype=torch.float32, device=dev), "free": torch.cuda.Event()}
            buf["free"].record()
            self._staged.append(buf)
        view = buf["buf"][: waves_host.numel()].view(waves_host.shape)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(buf["free"])
            view.copy_(waves_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        return {"view": view, "ready": ready, "slot": buf}

    def _resolve(self, waves):
        """Accept a tensor or a handle from stage(); returns (device tensor, slot to release or None)."""
        if isinstance(waves, dict):
            torch.cuda.current_stream().wait_event(waves["ready"])
            return waves["view"], waves["slot"]
        return waves, None

    @torch.no_grad()
    def logits(self, waves: torch.Tensor, n_samples: torch.Tensor, use_mask: bool = True):
        """waves (B, Nmax) fp32, n_samples (B,) -> logits (B, T', V), encoder lengths (B,) = mel frames // 4."""
        waves, slot = self._resolve(waves)
        feats, frames = self.preprocessor.extract_features_batch(waves, n_samples)
        if slot is not None:
            slot["free"].record()  # the front-end has consumed the staged waveforms
        out = self.model(feats, frames if use_mask else None)
        return out, torch.div(frames, 4, rounding_mode="floor")

    @torch.no_grad()
    def transcribe_ids(self, waves: torch.Tensor, n_samples: torch.Tensor, use_mask: bool = True) -> List[List[int]]:
        out, lengths = self.logits(waves, n_samples, use_mask)
        return self.decoder.decode_ids_batch(out, lengths if use_mask else None)

    @torch.no_grad()
    def submit(self, waves, n_samples: torch.Tensor, use_mask: bool = True):
        """Enqueue front-end, encoder, argmax/collapse and the device -> host copy of the token ids without waiting;
        `.result()` of the returned job blocks on that copy only.  Submitting batch i+1 before asking for the result of
        batch i keeps the GPU busy while the host turns ids into Python lists."""
        out, lengths = self.logits(waves, n_samples, use_mask)
        lens = lengths.to(device=out.device, dtype=torch.int64) if use_mask else None
        _, tokens, out_len = L.argmax_collapse(out, lens, blank=self.decoder.blank_id)
        host_tok, host_len = self._host_pair(tokens, out_len)
        host_tok.copy_(tokens, non_blocking=True)
        host_len.copy_(out_len, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        return _Job(host_tok, host_len, done, self._host_free)

    def _host_pair(self, tokens, out_len):
        """Pinned result buffers are pooled (cudaHostAlloc per batch would serialise the pipeline)."""
        for i, (t, l) in enumerate(self._host_free):
            if t.shape == tokens.shape and l.shape == out_len.shape:
                return self._host_free.pop(i)
        return (torch.empty(tokens.shape, dtype=tokens.dtype).pin_memory(),
                torch.empty(out_len.shape, dtype=out_len.dtype).pin_memory())

    def transcribe(self, waves: torch.Tensor, n_samples: torch.Tensor):
        ids = self.transcribe_ids(waves, n_samples)
        tok = self.decoder.tokenizer
        return [tok.decode(i) if tok is not None and hasattr(tok, "decode") else i for i in ids]
