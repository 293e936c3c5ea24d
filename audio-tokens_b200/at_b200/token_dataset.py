"""Token consumer served from HBM: the device-resident counterpart of the reference's ``TokenizedSpecDataset`` +
``collate_fn`` + ``DataLoaderCreator`` (datasets/tokenized_spec_dataset.py:13-76, datasets/data_loader_creator.py:17-34 of
danavery/audio-tokens) -- the step right after the hot path.

The reference opens one ``.npy`` token file per ``__getitem__`` and pads on the host for every batch.  Here the tokens of a
split live in ONE flat device array (straight from SpecTokenizer / HotPath, or read once from the token files) and a batch is
assembled by one kernel launch (at_tokens_collate + at_tokens_multihot).

* ``TokenStore``             flat tokens + per-clip offsets (+ CSR of label indices) on the device.
* ``TokenizedSpecDataset``   same constructor / ``__len__`` / ``__getitem__`` / ``collate_fn`` contract as the reference class;
                             ``__getitem__`` returns ``(seq, {"labels": labels})`` with CUDA tensors.
* ``DeviceBatchLoader``      iterable of ``(sequences, {"attention_masks", "labels"})`` like the reference's DataLoader over
                             ``collate_fn`` (shuffle with a seeded torch generator), batches assembled on the device.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import _lib


class TokenStore:
    """All token sequences of a split in one device array."""

    def __init__(self, tokens, offsets, names=None, label_ids=None, label_offsets=None):
        import torch

        _lib.require_cuda()
        assert tokens.is_cuda and tokens.dtype in (torch.int32, torch.int64) and tokens.dim() == 1
        self.tokens = tokens.contiguous()
        self.offsets = torch.as_tensor(offsets, dtype=torch.int64).cuda().contiguous()
        self.n = self.offsets.numel() - 1
        self.names = list(names) if names is not None else [str(i) for i in range(self.n)]
        self.label_ids = label_ids
        self.label_offsets = label_offsets
        self.lib = _lib.load()
        self._maxlen = torch.zeros(1, dtype=torch.int32, device="cuda")

    @classmethod
    def from_files(cls, files, threads: int = 16):
        """Token files (int64 (T,) .npy, spec_tokenizer.py:84-86) -> store; one bulk upload."""
        import torch
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=threads) as pool:
            arrs = list(pool.map(np.load, files))
        lens = np.array([len(a) for a in arrs], dtype=np.int64)
        off = np.zeros(len(arrs) + 1, dtype=np.int64)
        off[1:] = np.cumsum(lens)
        flat = np.concatenate(arrs).astype(np.int64) if arrs else np.zeros(0, dtype=np.int64)
        names = [os.path.splitext(os.path.basename(str(f)))[0] for f in files]
        return cls(torch.from_numpy(flat).cuda(), off, names)

    def set_labels(self, label_lists):
        """label_lists[i]: class indices of clip i (AudiosetMetadataProcessor.ytid_labels[ytid])."""
        import torch

        lens = np.array([len(l) for l in label_lists], dtype=np.int64)
        off = np.zeros(self.n + 1, dtype=np.int64)
        off[1:] = np.cumsum(lens)
        ids = np.concatenate([np.asarray(l, dtype=np.int32) for l in label_lists]) if lens.sum() else np.zeros(0, np.int32)
        self.label_ids = torch.from_numpy(ids.astype(np.int32)).cuda()
        if self.label_ids.numel() == 0:
            self.label_ids = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.label_offsets = torch.from_numpy(off).cuda()

    def sequence(self, i: int):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.tokens[a:b]

    def collate(self, idx, num_classes: int | None = None, reference_masks: bool = True, t_max: int | None = None):
        """idx: int64 CUDA tensor (B,) of clip ids -> (sequences (B, T_max) int64, attention_masks (B, T_max) float,
        labels (B, num_classes) float or None).  reference_masks=True reproduces the reference's all-ones masks."""
        import torch

        idx = idx.to(device="cuda", dtype=torch.int64).contiguous()
        B = idx.numel()
        if t_max is None:
            _lib.check(self.lib.at_tokens_batch_max_len(_lib.ptr(self.offsets), _lib.ptr(idx), B, _lib.ptr(self._maxlen),
                                                        _lib.stream_ptr()))
            t_max = int(self._maxlen.item())
        seqs = torch.empty((B, t_max), dtype=torch.int64, device="cuda")
        masks = torch.empty((B, t_max), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.at_tokens_collate(_lib.ptr(self.tokens), self.tokens.element_size(), _lib.ptr(self.offsets),
                                              _lib.ptr(idx), B, t_max, int(reference_masks), _lib.ptr(seqs), _lib.ptr(masks),
                                              _lib.stream_ptr()))
        labels = None
        if num_classes is not None and self.label_offsets is not None:
            labels = torch.empty((B, num_classes), dtype=torch.float32, device="cuda")
            _lib.check(self.lib.at_tokens_multihot(_lib.ptr(self.label_ids), _lib.ptr(self.label_offsets), _lib.ptr(idx), B,
                                                   num_classes, _lib.ptr(labels), _lib.stream_ptr()))
        return seqs, masks, labels


class TokenizedSpecDataset:
    """Mirror of the reference class (same constructor arguments and item / batch contract) over a TokenStore.

    config: needs split_file, tokenized_train_dir / tokenized_val_dir, num_classes; data_manager: object with
    ``ytid_labels`` (dict ytid -> list of class indices), as AudiosetMetadataProcessor provides.  ``store`` lets a caller
    hand over tokens that are already on the device (no files read)."""

    def __init__(self, config, data_manager, split: str = "train", store: TokenStore | None = None):
        self.config, self.data_manager, self.split = config, data_manager, split
        with open(config.split_file, "r") as f:
            self.ytids = json.load(f)[split]
        base = config.tokenized_train_dir if split == "train" else config.tokenized_val_dir
        if store is None:
            files = [os.path.join(base, f"{y}.npy") for y in self.ytids]
            self.tokenized_spec_files = [f for f in files if os.path.exists(f)]   # missing clips are skipped silently
            store = TokenStore.from_files(self.tokenized_spec_files)
        else:
            keep = set(self.ytids)
            assert all(n in keep for n in store.names), "store holds clips outside this split"
            self.tokenized_spec_files = [os.path.join(base, f"{n}.npy") for n in store.names]
        self.store = store
        self.store.set_labels([data_manager.ytid_labels[n] for n in store.names])

    def __len__(self):
        return self.store.n

    def __getitem__(self, idx: int):
        import torch

        seq = self.store.sequence(idx)
        labels = torch.zeros(self.config.num_classes, dtype=torch.float, device="cuda")
        li = self.data_manager.ytid_labels[self.store.names[idx]]
        if len(li):
            labels[torch.as_tensor(li, dtype=torch.int64, device="cuda")] = 1.0
        return seq, {"labels": labels}

    @staticmethod
    def collate_fn(batch):
        """The reference's collate_fn verbatim in meaning (host-side composition of already-fetched items; the fast path
        is DeviceBatchLoader): pad with 0, .long(), masks built from the PADDED matrix (all ones), stacked labels."""
        import torch
        from torch.nn.utils.rnn import pad_sequence

        sequences, metadata = zip(*batch)
        labels = [item["labels"] for item in metadata]
        sequences = pad_sequence(sequences, batch_first=True, padding_value=0).long()
        attention_masks = torch.ones_like(sequences).float()
        return sequences, {"attention_masks": attention_masks, "labels": torch.stack(labels).float()}


class DeviceBatchLoader:
    """DataLoader(dataset, batch_size, shuffle, collate_fn=dataset.collate_fn) of data_loader_creator.py:20-33, with the
    batch assembled on the device.  shuffle uses torch.randperm on a seeded generator (torch's RandomSampler does the
    same); drop_last False like the reference."""

    def __init__(self, dataset: TokenizedSpecDataset, batch_size: int, shuffle: bool = False, seed: int | None = None,
                 reference_masks: bool = True):
        self.ds, self.bs, self.shuffle, self.reference_masks = dataset, int(batch_size), shuffle, reference_masks
        import torch

        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)

    def __len__(self):
        return (len(self.ds) + self.bs - 1) // self.bs

    def __iter__(self):
        import torch

        n = len(self.ds)
        order = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        order = order.cuda()
        for i in range(0, n, self.bs):
            seqs, masks, labels = self.ds.store.collate(order[i:i + self.bs], self.ds.config.num_classes, self.reference_masks)
            yield seqs, {"attention_masks": masks, "labels": labels}
