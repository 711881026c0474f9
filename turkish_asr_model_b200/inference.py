"""Batched long-form inference up to token ids (BASELINE configs[3]).

The reference transcribes one file at a time without a padding mask (inference.py:101-148: preprocessor ->
model(features) -> argmax -> tokenizer.ctc_decode).  Here a whole batch of waveforms goes through the GPU
front-end, the encoder with the key-padding mask of the training path, and the argmax/collapse kernel."""
from typing import List, Optional

import torch

from .data.preprocessing import AudioPreprocessor
from .utils.decoding import GreedyDecoder


class BatchedInference:
    def __init__(self, model, preprocessor: Optional[AudioPreprocessor] = None, tokenizer=None, blank_id: int = 0):
        self.model = model.eval()
        self.preprocessor = preprocessor if preprocessor is not None else AudioPreprocessor(device="cuda")
        self.decoder = GreedyDecoder(tokenizer, blank_id=blank_id)

    @torch.no_grad()
    def logits(self, waves: torch.Tensor, n_samples: torch.Tensor, use_mask: bool = True):
        """waves (B, Nmax) fp32, n_samples (B,) -> logits (B, T', V), encoder lengths (B,) = mel frames // 4."""
        feats, frames = self.preprocessor.extract_features_batch(waves, n_samples)
        out = self.model(feats, frames if use_mask else None)
        return out, torch.div(frames, 4, rounding_mode="floor")

    @torch.no_grad()
    def transcribe_ids(self, waves: torch.Tensor, n_samples: torch.Tensor, use_mask: bool = True) -> List[List[int]]:
        out, lengths = self.logits(waves, n_samples, use_mask)
        return self.decoder.decode_ids_batch(out, lengths if use_mask else None)

    def transcribe(self, waves: torch.Tensor, n_samples: torch.Tensor):
        ids = self.transcribe_ids(waves, n_samples)
        tok = self.decoder.tokenizer
        return [tok.decode(i) if tok is not None and hasattr(tok, "decode") else i for i in ids]
