"""Drop-in mirror of the hot-path pieces of reference data/dataset.py: BucketingSampler (:123-172) with the
new rank/world sharding for data parallelism, and collate_fn (:283-312)."""
import random
from typing import List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Sampler


class BucketingSampler(Sampler):
    """Sort by length proxy, chunk into buckets of `batch_size`, shuffle the buckets, flatten
    (reference data/dataset.py:149-167).  With world_size == 1 the flat order is bit-identical to the
    reference (global `random.shuffle`, or a private Random(seed + epoch) stream when `seed` is given).

    Data parallel (new): the reference algorithm runs with bucket size world_size*batch_size;
    rank r takes elements [r*B, (r+1)*B) of every global bucket, so the union over ranks equals the reference
    run at batch W*B and all ranks see similar lengths in the same step.  A final partial global bucket is
    dropped so that every rank performs the same number of steps.

    `data_source` is either the reference's dataset (has .file_pairs -> os.path.getsize) or any sequence of
    integer lengths."""

    def __init__(self, data_source, batch_size: int, shuffle: bool = True, drop_last: bool = False, *, rank: int = 0,
                 world_size: int = 1, seed: Optional[int] = None, lengths: Optional[Sequence[int]] = None):
        self.data_source = data_source
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.rank = rank
        self.world_size = world_size
        self.seed = seed
        self.epoch = 0
        if lengths is not None:
            self.lengths = list(lengths)
        elif hasattr(data_source, "file_pairs"):
            import os
            self.lengths = []
            for wav_path, _ in data_source.file_pairs:
                try:
                    self.lengths.append(os.path.getsize(wav_path))
                except OSError:
                    self.lengths.append(0)
        else:
            self.lengths = [int(v) for v in data_source]

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def _flat_order(self, bucket):
        indices = sorted(range(len(self.lengths)), key=lambda i: self.lengths[i])
        batches = []
        for i in range(0, len(indices), bucket):
            batch = indices[i:i + bucket]
            if len(batch) == bucket or not self.drop_last:
                batches.append(batch)
        if self.shuffle:
            if self.seed is None:
                random.shuffle(batches)
            else:
                random.Random(self.seed + self.epoch).shuffle(batches)
        return batches

    def __iter__(self):
        if self.world_size == 1:
            for batch in self._flat_order(self.batch_size):
                yield from batch
            return
        gb = self.batch_size * self.world_size
        for batch in self._flat_order(gb):
            if len(batch) < gb:
                continue
            yield from batch[self.rank * self.batch_size:(self.rank + 1) * self.batch_size]

    def __len__(self) -> int:
        n = len(self.lengths)
        if self.world_size == 1:
            return (n // self.batch_size) * self.batch_size if self.drop_last else n
        return (n // (self.batch_size * self.world_size)) * self.batch_size


def collate_fn(batch: List[Tuple[torch.Tensor, torch.Tensor]]):
    """reference data/dataset.py:283-312: zero-pad features to Tmax and targets to Smax, return true lengths."""
    batch = [item for item in batch if item is not None and item[0] is not None]
    if len(batch) == 0:
        return None, None, None, None
    features, targets = zip(*batch)
    input_lengths = torch.LongTensor([f.size(0) for f in features])
    target_lengths = torch.LongTensor([len(t) for t in targets])
    features_padded = torch.nn.utils.rnn.pad_sequence(features, batch_first=True)
    targets_padded = torch.nn.utils.rnn.pad_sequence(targets, batch_first=True, padding_value=0)
    return features_padded, targets_padded, input_lengths, target_lengths
