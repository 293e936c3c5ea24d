"""The token consumer: device-resident TokenizedSpecDataset / collate against the reference's
expressions (datasets/tokenized_spec_dataset.py:52-76), restated here on the CPU from the token files."""
import json
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_item(path, label_indices, num_classes):
    import torch

    seq = torch.tensor(np.load(path))                       # :54
    labels = torch.zeros(num_classes, dtype=torch.float)    # :59-60
    labels[label_indices] = 1.0
    return seq, {"labels": labels}


def _reference_collate(batch):
    import torch
    from torch.nn.utils.rnn import pad_sequence

    sequences, metadata = zip(*batch)                        # :66-76
    labels = [item["labels"] for item in metadata]
    sequences = pad_sequence(sequences, batch_first=True, padding_value=0).long()
    attention_masks = pad_sequence([torch.ones_like(seq) for seq in sequences], batch_first=True, padding_value=0).float()
    return sequences, {"attention_masks": attention_masks, "labels": torch.stack(labels).float()}


def test_device_dataset_and_loader_equal_the_reference_expressions(tmp_path):
    import torch
    from at_b200.token_dataset import DeviceBatchLoader, TokenizedSpecDataset, TokenStore

    rng = np.random.default_rng(3)
    ytids = [f"yt{i:03d}" for i in range(37)]
    tok_dir = tmp_path / "tokenized_audio" / "train"
    tok_dir.mkdir(parents=True)
    lens = rng.integers(1, 90, size=len(ytids))
    present = [y for i, y in enumerate(ytids) if i % 9 != 4]      # some clips have no token file: skipped silently
    for y, n in zip(ytids, lens):
        if y in present:
            np.save(tok_dir / f"{y}.npy", rng.integers(0, 500, size=n).astype(np.int64))
    split = tmp_path / "split.json"
    json.dump({"train": ytids, "validation": []}, open(split, "w"))
    num_classes = 23
    labels = {y: sorted(rng.choice(num_classes, size=rng.integers(0, 4), replace=False).tolist()) for y in ytids}
    cfg = types.SimpleNamespace(split_file=str(split), tokenized_train_dir=str(tok_dir), tokenized_val_dir=str(tok_dir),
                                num_classes=num_classes, training_batch_size=8)
    dm = types.SimpleNamespace(ytid_labels=labels)
    ds = TokenizedSpecDataset(cfg, dm, "train")
    assert len(ds) == len(present)
    # items
    for i in (0, 5, len(ds) - 1):
        seq, meta = ds[i]
        rseq, rmeta = _reference_item(ds.tokenized_spec_files[i], labels[present[i]], num_classes)
        assert torch.equal(seq.cpu(), rseq) and torch.equal(meta["labels"].cpu(), rmeta["labels"])
    # the class's own collate_fn on device items == the reference's on host items
    batch = [ds[i] for i in range(6)]
    s, m = TokenizedSpecDataset.collate_fn(batch)
    rs, rm = _reference_collate([_reference_item(ds.tokenized_spec_files[i], labels[present[i]], num_classes) for i in range(6)])
    assert torch.equal(s.cpu(), rs) and torch.equal(m["attention_masks"].cpu(), rm["attention_masks"])
    assert torch.equal(m["labels"].cpu(), rm["labels"])
    # the device loader, unshuffled: every batch equals the reference DataLoader's batch
    loader = DeviceBatchLoader(ds, cfg.training_batch_size, shuffle=False)
    seen = 0
    for bi, (seqs, meta) in enumerate(loader):
        ids = range(bi * 8, min(len(ds), bi * 8 + 8))
        rs, rm = _reference_collate([_reference_item(ds.tokenized_spec_files[i], labels[present[i]], num_classes) for i in ids])
        assert seqs.dtype == torch.int64 and torch.equal(seqs.cpu(), rs)
        assert torch.equal(meta["attention_masks"].cpu(), rm["attention_masks"])
        assert torch.equal(meta["labels"].cpu(), rm["labels"])
        seen += seqs.shape[0]
    assert seen == len(ds) and len(loader) == (len(ds) + 7) // 8
    # shuffled: a permutation of the clips, true (length-aware) masks on request
    loader = DeviceBatchLoader(ds, 5, shuffle=True, seed=11, reference_masks=False)
    total = 0
    for seqs, meta in loader:
        lens_b = meta["attention_masks"].sum(1).long()
        assert int(lens_b.max()) == seqs.shape[1]
        assert bool(((seqs != 0) <= (meta["attention_masks"] > 0)).all())
        total += int(lens_b.sum())
    assert total == int(sum(len(np.load(f)) for f in ds.tokenized_spec_files))
    # a store handed over from the tokenizer (int32 labels straight from the search kernel)
    flat = torch.from_numpy(np.concatenate([np.load(f) for f in ds.tokenized_spec_files]).astype(np.int32)).cuda()
    st = TokenStore(flat, ds.store.offsets.cpu().numpy(), ds.store.names)
    ds2 = TokenizedSpecDataset(cfg, dm, "train", store=st)
    a = ds.store.collate(torch.arange(7).cuda(), num_classes)
    b = ds2.store.collate(torch.arange(7).cuda(), num_classes)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
