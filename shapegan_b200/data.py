"""Voxel data path: `VoxelDataset` keeps the reference's class surface (datasets.py:7-42: ctor, `glob`, `from_split`,
`__getitem__` = np.load -> clamp_(-c, c) -> /= c on the CPU, usable with torch's DataLoader exactly like the reference), and
`VoxelBatchStream` is the B200-native ingest for the training loops: raw .npy grids are read by a background thread straight into
PINNED staging buffers, copied to the device on a side stream and clamped / rescaled there by sg_voxel_ingest (bit-identical to
`__getitem__`), double-buffered so batch i+1 loads and uploads while batch i trains."""
import os
import queue
import threading

import numpy as np
import torch
from torch.utils.data import Dataset

from . import raw


class VoxelDataset(Dataset):
    def __init__(self, files, clamp=0.1, rescale_sdf=True):
        self.files = files
        self.clamp = clamp
        self.rescale_sdf = rescale_sdf

    def __len__(self):
        return len(self.files)

    def __getitem__(self, index):
        result = torch.from_numpy(np.load(self.files[index]))
        if self.clamp is not None:
            result.clamp_(-self.clamp, self.clamp)
            if self.rescale_sdf:
                result /= self.clamp
        return result

    @staticmethod
    def glob(pattern):
        import glob
        files = glob.glob(pattern, recursive=True)
        if len(files) == 0:
            raise Exception('No files found for glob pattern {:s}.'.format(pattern))
        return VoxelDataset(sorted(files))

    @staticmethod
    def from_split(pattern, split_file_name):
        with open(split_file_name, 'r') as split_file:
            ids = split_file.readlines()
        files = [pattern.format(id.strip()) for id in ids]
        return VoxelDataset([file for file in files if os.path.exists(file)])

    def stream(self, batch_size, device='cuda', shuffle=True, drop_last=False, seed=None, depth=2):
        """B200-native replacement of `DataLoader(dataset, shuffle=True, batch_size=B, num_workers=8)`"""
        return VoxelBatchStream(self, batch_size, device, shuffle, drop_last, seed, depth)


class VoxelBatchStream:
    """Iterating yields device tensors [b, R, R, R] fp32, already clamped / rescaled, one epoch per iteration (like a DataLoader).

    Pipeline per batch: loader thread  np.load -> pinned staging[slot]      (no intermediate torch tensors, no collate copy)
                        side stream    staging[slot] -> raw[slot] (H2D), sg_voxel_ingest(raw[slot]) -> batch[slot]
                        consumer       waits on the slot's event only: upload and ingest of batch i+1 overlap the step on batch i."""

    def __init__(self, dataset, batch_size, device='cuda', shuffle=True, drop_last=False, seed=None, depth=2):
        if not torch.cuda.is_available():
            raise RuntimeError('VoxelBatchStream: needs a CUDA device (use VoxelDataset with a DataLoader on CPU hosts)')
        self.ds, self.b, self.device = dataset, batch_size, torch.device(device)
        self.shuffle, self.drop_last, self.depth = shuffle, drop_last, max(2, depth)
        self.rng = np.random.default_rng(seed)
        shape = tuple(np.load(dataset.files[0], mmap_mode='r').shape)
        self.shape = shape
        self.staging = [torch.empty((batch_size,) + shape, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        self.raw = [torch.empty((batch_size,) + shape, dtype=torch.float32, device=self.device) for _ in range(self.depth)]
        self.out = [torch.empty((batch_size,) + shape, dtype=torch.float32, device=self.device) for _ in range(self.depth)]
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.consumed = [torch.cuda.Event() for _ in range(self.depth)]

    def __len__(self):
        n = len(self.ds)
        return n // self.b if self.drop_last else (n + self.b - 1) // self.b

    def _batches(self):
        order = self.rng.permutation(len(self.ds)) if self.shuffle else np.arange(len(self.ds))
        for i in range(0, len(order), self.b):
            idx = order[i:i + self.b]
            if len(idx) < self.b and self.drop_last:
                return
            yield idx

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth - 1)
        free = queue.Queue()
        for s in range(self.depth):
            free.put(s)

        def loader():
            try:
                for idx in self._batches():
                    slot = free.get()
                    buf = self.staging[slot].numpy()
                    for j, k in enumerate(idx):
                        buf[j] = np.load(self.ds.files[k])
                    q.put((slot, len(idx)))
            except Exception as e:          # surfaced in the consumer
                q.put(e)
            q.put(None)
        t = threading.Thread(target=loader, daemon=True)
        t.start()
        clamp = self.ds.clamp
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            slot, count = item
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.consumed[slot])          # the previous user of this slot's device buffers is done
                self.raw[slot][:count].copy_(self.staging[slot][:count], non_blocking=True)
                if clamp is not None:
                    raw.voxel_ingest(self.raw[slot][:count], clamp, self.ds.rescale_sdf, out=self.out[slot][:count])
                else:
                    self.out[slot][:count].copy_(self.raw[slot][:count])
                self.ready[slot].record(self.stream)
            self.ready[slot].synchronize()                           # staging[slot] may be refilled once the H2D copy has run
            free.put(slot)
            torch.cuda.current_stream().wait_event(self.ready[slot])
            yield self.out[slot][:count]
            self.consumed[slot].record(torch.cuda.current_stream())
        t.join()
