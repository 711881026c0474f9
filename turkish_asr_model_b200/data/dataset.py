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
        """Length-sorted index list cut into buckets of `bucket`, optionally shuffled as whole buckets.  The stable sort
        and the single list shuffle consume the random stream exactly like the reference, so orders are bit-identical."""
        n = len(self.lengths)
        by_length = sorted(range(n), key=self.lengths.__getitem__)
        buckets = [by_length[lo: lo + bucket] for lo in range(0, n, bucket)]
        if self.drop_last and buckets and len(buckets[-1]) < bucket:
            buckets.pop()
        if self.shuffle:
            rng = random if self.seed is None else random.Random(self.seed + self.epoch)
            rng.shuffle(buckets)
        return buckets

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
    """Batch assembly with the contract of reference data/dataset.py:283-312: items that failed to load (None) are
    dropped; an empty batch yields four Nones; otherwise features (T_b, F) are zero-padded to the longest T and targets
    (S_b,) to the longest S with id 0, and the true lengths come back as int64 vectors."""
    kept = [(f, t) for item in batch if item is not None for (f, t) in [item] if f is not None]
    if not kept:
        return None, None, None, None
    n = len(kept)
    frames = [int(f.shape[0]) for f, _ in kept]
    labels = [int(t.shape[0]) for _, t in kept]
    f0, t0 = kept[0]
    feats = f0.new_zeros((n, max(frames)) + tuple(f0.shape[1:]))
    targets = t0.new_zeros((n, max(labels)) + tuple(t0.shape[1:]))
    for b, (f, t) in enumerate(kept):
        feats[b, : frames[b]] = f
        targets[b, : labels[b]] = t
    return feats, targets, torch.tensor(frames, dtype=torch.int64), torch.tensor(labels, dtype=torch.int64)
