"""Batched spectrogram-file reader against the reference's expression (np.load(f).T per file + np.concatenate)."""
import logging
import types

import numpy as np
import pytest

from at_b200.npyio import load_spec_batch


def _write_files(tmp_path, n=23, d=64, fortran=True, seed=0):
    rng = np.random.default_rng(seed)
    files = []
    for i in range(n):
        t = int(rng.integers(0, 40)) if i % 7 else 0       # a few empty clips
        frames = rng.random((t, d), dtype=np.float32)        # frame-major, as the spectrogram stage holds it
        arr = frames.T if fortran else np.ascontiguousarray(frames.T)   # (d, T); .T of a C array saves fortran_order=True
        p = tmp_path / f"clip{i:03d}.npy"
        np.save(p, arr)
        files.append(p)
    return files


@pytest.mark.parametrize("fortran", [True, False])
@pytest.mark.parametrize("threads", [1, 8])
def test_batch_reader_equals_reference_expression(tmp_path, fortran, threads):
    files = _write_files(tmp_path, fortran=fortran)
    want = np.concatenate([np.load(f).T for f in files], axis=0).astype(np.float32)
    got, lengths = load_spec_batch(files, threads=threads)
    assert got.dtype == np.float32 and got.flags.c_contiguous
    assert np.array_equal(got, want)
    assert lengths == [np.load(f).shape[1] for f in files]
    assert load_spec_batch([])[1] == []


def test_batch_reader_rejects_mixed_widths(tmp_path):
    np.save(tmp_path / "a.npy", np.zeros((64, 3), dtype=np.float32))
    np.save(tmp_path / "b.npy", np.zeros((32, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        load_spec_batch([tmp_path / "a.npy", tmp_path / "b.npy"])


def test_cluster_creator_batches_equal_reference_expression(tmp_path):
    """ClusterCreator._batch_generator (processors/cluster_creator.py:83-102) without a device: the stage object is
    built around its config only."""
    from processors.cluster_creator import ClusterCreator

    train = tmp_path / "train"
    train.mkdir()
    files = sorted(_write_files(train, n=11, seed=3))
    cc = object.__new__(ClusterCreator)
    cc.config = types.SimpleNamespace(source_spec_path=str(tmp_path), sort_files=True)
    cc.logger = logging.getLogger("test")
    batches = list(cc._batch_generator(4))
    assert len(batches) == 3
    for i, b in enumerate(batches):
        want = np.concatenate([np.load(f).T for f in files[4 * i:4 * i + 4]], axis=0).astype(np.float32)
        assert np.array_equal(b, want)


def test_spec_tokenizer_process_batch_splits_tokens_per_file(tmp_path):
    """SpecTokenizer.process_batch (processors/spec_tokenizer.py:66-90) with a stand-in index: every file gets the int64
    tokens of its own frames, in order, and the flat list is returned."""
    from processors.spec_tokenizer import SpecTokenizer

    src = tmp_path / "src"
    dst = tmp_path / "dst"
    src.mkdir(), dst.mkdir()
    files = sorted(_write_files(src, n=9, seed=5))

    class FakeIndex:
        def search(self, x, k, l2norm_rows=False, want_dist=True):
            assert k == 1 and l2norm_rows and x.dtype == np.float32 and x.shape[1] == 64
            lab = np.arange(x.shape[0], dtype=np.int64)[:, None]   # token = global frame number: easy to check
            return np.zeros((x.shape[0], 1), dtype=np.float32), lab

    import types

    st = object.__new__(SpecTokenizer)
    st.index = FakeIndex()
    st.logger = logging.getLogger("test")
    st.config = types.SimpleNamespace(use_convolution=False)   # process_batch reads it like the reference's (:70)
    flat = st.process_batch(files, dst)
    lengths = [np.load(f).shape[1] for f in files]
    assert flat == list(range(sum(lengths)))
    pos = 0
    for f, n in zip(files, lengths):
        tok = np.load(dst / f"{f.stem}.npy")
        assert tok.dtype == np.int64 and tok.shape == (n,)
        assert np.array_equal(tok, np.arange(pos, pos + n))
        pos += n
    assert st.process_batch([], dst) == []


def test_token_statistics_match_the_reference_expressions():
    """SpecTokenizer.token_statistics (the numbers of plot_token_distribution / analyze_zipf_and_tail, reference
    processors/spec_tokenizer.py:146-236) against the reference's own expressions: Counter, sorted, np.cumsum / searchsorted,
    scipy.stats.linregress over the middle 80 % of the log-log curve."""
    from collections import Counter

    from scipy import stats

    from processors.spec_tokenizer import SpecTokenizer

    rng = np.random.default_rng(3)
    k = 300
    all_tokens = rng.zipf(1.3, size=200000) % k   # long-tailed, a few unused ids
    counts = np.bincount(all_tokens, minlength=k)
    st = SpecTokenizer.token_statistics(counts)
    token_counts = Counter(all_tokens.tolist())
    sorted_counts = sorted(token_counts.items(), key=lambda x: x[1], reverse=True)
    tokens, frequencies = zip(*sorted_counts)
    assert list(st["frequencies"]) == list(frequencies) and st["total"] == sum(frequencies) and len(st["tokens"]) == len(tokens)
    cumulative_freq = np.cumsum(frequencies) / sum(frequencies)
    assert st["top_80_percent"] == np.searchsorted(cumulative_freq, 0.8) + 1
    ranks = np.arange(1, len(frequencies) + 1)
    a, b = int(0.1 * len(frequencies)), int(0.9 * len(frequencies))
    slope, _, r_value, _, _ = stats.linregress(np.log(ranks)[a:b], np.log(frequencies)[a:b])
    assert abs(st["slope"] - slope) < 1e-9 and abs(st["r_squared"] - r_value ** 2) < 1e-9
    tail_start = np.searchsorted(cumulative_freq, 0.8)
    assert st["tail_start"] == tail_start and abs(st["tail_proportion"] - (1 - tail_start / len(frequencies))) < 1e-12
