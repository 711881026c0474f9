"""BucketingSampler / collate oracle (pure Python).  TEST INFRASTRUCTURE ONLY.

data/dataset.py:149-167: stable sort of indices by file size, consecutive chunks of batch_size,
random.shuffle(batches) with the global Python RNG, flatten; DataLoader re-chunks the flat stream.
data/dataset.py:283-312: collate pads features/targets with zeros and returns the true lengths.
"""
import random


def bucketing_buckets(lengths, batch_size, shuffle=True, drop_last=False, rng=random):
    """The shuffled list of buckets (data/dataset.py:151-163)."""
    indices = sorted(range(len(lengths)), key=lambda i: lengths[i])
    batches = []
    for i in range(0, len(indices), batch_size):
        batch = indices[i:i + batch_size]
        if len(batch) == batch_size or not drop_last:
            batches.append(batch)
    if shuffle:
        rng.shuffle(batches)
    return batches


def bucketing_order(lengths, batch_size, shuffle=True, drop_last=False, rng=random):
    """Flat index stream the sampler yields (data/dataset.py:165-167)."""
    return [i for b in bucketing_buckets(lengths, batch_size, shuffle, drop_last, rng) for i in b]


def dataloader_batches(flat, batch_size):
    return [flat[i:i + batch_size] for i in range(0, len(flat), batch_size)]
