"""Drop-in mirror of reference utils/decoding.py:132-169 (GreedyDecoder) up to token ids: argmax, repeat
collapse and blank removal run in one CUDA kernel pair (tasr_argmax_collapse); turning ids into text is the
tokenizer's job (data/tokenizer.py needs the HF hub and is out of scope)."""
from typing import List, Optional

import torch

from .. import _lib as L


class GreedyDecoder:
    def __init__(self, tokenizer=None, blank_id: int = 0):
        self.tokenizer = tokenizer
        self.blank_id = blank_id

    def decode_ids_batch(self, logits: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> List[List[int]]:
        """logits (B, T, V) on CUDA -> collapsed token id lists (bit-exact vs argmax + ctc_decode filtering)."""
        if lengths is not None:
            lengths = lengths.to(device=logits.device, dtype=torch.int64)
        _, tokens, out_len = L.argmax_collapse(logits, lengths, blank=self.blank_id)
        tokens, out_len = tokens.cpu(), out_len.cpu()
        return [tokens[b, : int(out_len[b])].tolist() for b in range(tokens.shape[0])]

    def _text(self, ids):
        if self.tokenizer is None:
            return ids
        return self.tokenizer.decode(ids) if hasattr(self.tokenizer, "decode") else ids

    def decode(self, logits: torch.Tensor):
        """logits (T, V) -> text (or ids when no tokenizer is attached)."""
        return self._text(self.decode_ids_batch(logits.unsqueeze(0))[0])

    def decode_batch(self, logits: torch.Tensor):
        return [self._text(ids) for ids in self.decode_ids_batch(logits)]
