"""Input feed for the hot path.

The reference builds ``DataLoader(train_ds, batch_size, shuffle, num_workers=0, pin_memory)`` with no sampler hook
(data_loader.py:29-44,239) and copies each batch to the device synchronously inside the step (train.py:220-221).
At B200 throughput (>= 1400 samples/s x 1 MB fp32 per sample) that serialises the step behind the copy, and there is
no rank sharding for data-parallel runs.  Two small pieces fix both without touching the reference's Dataset:

* ``RankShardSampler`` -- rank r takes samples r, r + world, ... of a (seeded, per-epoch) permutation; plugs into
  ``DataLoader(dataset, sampler=...)``.
* ``DevicePrefetcher`` -- wraps any iterable of (specs, labels) host batches; stages them in pinned memory and
  issues the host->device copy of batch k+1 on a side stream while batch k computes (double buffered, event
  ordered, no host synchronisation).  ``bench.py``'s end-to-end leg uses it.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch
from torch.utils.data import Sampler


class RankShardSampler(Sampler[int]):
    def __init__(self, data_len: int, rank: int, world: int, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = False):
        if not (0 <= rank < world):
            raise ValueError(f"rank {rank} not in [0, {world})")
        self.n, self.rank, self.world = int(data_len), rank, world
        self.shuffle, self.seed, self.drop_last = shuffle, seed, drop_last
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _order(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).tolist()
        else:
            order = list(range(self.n))
        if self.drop_last:
            order = order[:self.n - self.n % self.world]
        else:                                   # pad by wrapping so that every rank sees the same number of samples
            pad = (-len(order)) % self.world
            order += order[:pad]
        return order

    def __iter__(self) -> Iterator[int]:
        return iter(self._order()[self.rank::self.world])

    def __len__(self) -> int:
        if self.drop_last:
            return self.n // self.world
        return (self.n + self.world - 1) // self.world


class DevicePrefetcher:
    """Iterate device-resident (specs, labels) batches with the H2D copy of the next batch overlapped."""

    def __init__(self, loader: Iterable, device, depth: int = 2):
        self.loader = loader
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher feeds a CUDA device")
        self.depth = max(2, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch, slot):
        """Copy one host batch into slot's pinned staging buffers and start its async H2D copy."""
        pinned, dev, ready, free = slot
        ready.synchronize()                      # host: the previous copy out of this slot's pinned buffers is done
        self.stream.wait_event(free)             # device: the consumer's kernels on its previous batch are done
        out = []
        for i, t in enumerate(batch):
            t = torch.as_tensor(t)
            if i >= len(pinned) or pinned[i] is None or pinned[i].shape != t.shape or pinned[i].dtype != t.dtype:
                while len(pinned) <= i:
                    pinned.append(None)
                    dev.append(None)
                pinned[i] = t if t.is_pinned() else torch.empty(t.shape, dtype=t.dtype).pin_memory()
                dev[i] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            if pinned[i] is not t:
                if t.is_pinned():
                    pinned[i] = t
                else:
                    pinned[i].copy_(t)
            out.append(dev[i])
        with torch.cuda.stream(self.stream):
            for i in range(len(out)):
                dev[i].copy_(pinned[i], non_blocking=True)
                self.h2d_bytes += pinned[i].numel() * pinned[i].element_size()
            ready.record(self.stream)
        return tuple(out)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        slots = [([], [], torch.cuda.Event(), torch.cuda.Event()) for _ in range(self.depth)]
        it = iter(self.loader)
        queue = []
        k = 0
        try:
            for _ in range(self.depth - 1):
                queue.append((self._stage(next(it), slots[k % self.depth]), slots[k % self.depth]))
                k += 1
        except StopIteration:
            pass
        cur = torch.cuda.current_stream(self.device)
        while queue:
            tensors, slot = queue.pop(0)
            try:                                  # keep the pipe full before handing the current batch out
                queue.append((self._stage(next(it), slots[k % self.depth]), slots[k % self.depth]))
                k += 1
            except StopIteration:
                pass
            cur.wait_event(slot[2])
            yield tensors
            slot[3].record(cur)                   # consumer's kernels on this batch are enqueued
