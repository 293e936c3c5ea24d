"""Stage 3 drop-in: same class and outputs as the reference's processors/spec_tokenizer.py; the CPU
faiss.IndexFlatL2 search is replaced by at_index_search with the row normalisation fused into the kernel."""
import logging
import shutil
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm

import at_b200
import at_b200.faiss_compat as faiss
from at_b200 import _lib
from at_b200.npyio import load_spec_batch


def _set_seed(seed=42):
    """utils/set_seed.py of the reference (the conv layer's default initialisation draws from torch's CPU generator)."""
    import random

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


class SpecTokenizer:
    def __init__(self, config):
        self.config = config
        _set_seed(self.config.random_seed)
        self.logger = logging.getLogger()
        if not torch.cuda.is_available():
            raise RuntimeError("SpecTokenizer (B200 build) needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda")
        self.source_path = Path(self.config.source_spec_path)
        self.dest_tokenized_path = Path(self.config.dest_tokenized_path)
        self.centroid_path = Path(self.config.centroids_path)
        self.index = self.load_centroid_index()
        self.token_counts = None
        if self.config.use_convolution:
            self.conv_weight, self.conv_bias = self.create_convolution_layer()

    def run(self):
        for split in ["train", "validation"]:
            source_spec_dir = self.source_path / split
            tokenized_dir = self.dest_tokenized_path / split
            self.setup_output_directory(tokenized_dir)
            self.logger.info(f"Tokenizing {split} set: {source_spec_dir} --> {tokenized_dir}")
            all_tokens = self.tokenize_directory(source_spec_dir, tokenized_dir)
            if split == "train":
                self.analyze_tokens(all_tokens)
                self.plot_token_distribution(all_tokens)

    def tokenize_directory(self, source_dir: Path, tokenized_dir: Path):
        all_tokens = []
        spec_files = list(source_dir.glob("*.npy"))
        if getattr(self.config, "sort_files", False):
            spec_files = sorted(spec_files)
        bs = self.config.tokenizer_batch_size
        for i in tqdm(range(0, len(spec_files), bs)):
            all_tokens.extend(self.process_batch(spec_files[i: i + bs], tokenized_dir))
        return all_tokens

    def process_batch(self, batch_files, tokenized_dir: Path):
        if not batch_files:
            return []
        batch_data, lengths = load_spec_batch(batch_files)   # the reference's per-file np.load(f).T + concatenate
        if batch_data.size == 0:
            return []
        if self.config.use_convolution:
            # conv expansion + normalize_vectors on the device, then the wide-row exact search
            wide = at_b200.row_l2norm(self.apply_convolution(batch_data))
            _, tokens = self.index.search(wide, 1, want_dist=False)   # the reference discards D (:77)
        else:
            # normalize_vectors + index.search(x, 1) in one kernel; int64 labels like faiss
            _, tokens = self.index.search(batch_data, 1, l2norm_rows=True, want_dist=False)   # the reference discards D (:77)
        tokens = np.squeeze(tokens, 1)
        start = 0
        for spec_file, n_frames in zip(batch_files, lengths):
            end = start + n_frames
            np.save(tokenized_dir / f"{spec_file.stem}.npy", tokens[start:end])
            start = end
        return tokens.tolist()

    def apply_convolution(self, batch):
        """(n, n_mels) numpy / CUDA tensor -> (n, n_mels * num_kernels) CUDA tensor (reference :92-104); None for an empty
        batch like the reference."""
        if len(batch) == 0:
            self.logger.warning("Received empty batch for convolution")
            return None
        if not torch.is_tensor(batch):
            batch = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32)).to(self.device)
        return at_b200.conv_expand(batch.contiguous(), self.conv_weight, self.conv_bias)

    def create_convolution_layer(self):
        """(weight, bias) of the reference's seeded, never-trained nn.Conv1d (:115-121) as device tensors."""
        return at_b200.make_conv_layer(self.config)

    @staticmethod
    def normalize_vectors(vectors):
        v = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).cuda()
        return at_b200.row_l2norm(v).cpu().numpy()

    def setup_output_directory(self, tokenized_dir):
        shutil.rmtree(tokenized_dir, ignore_errors=True)
        tokenized_dir.mkdir(parents=True)

    def load_centroid_index(self):
        centroids = np.load(self.centroid_path)
        index = faiss.IndexFlatL2(centroids.shape[1])
        index.add(centroids)
        return index

    def analyze_tokens(self, all_tokens):
        """Token histogram (the reference builds a Counter over a Python list; here a device bincount) and the
        same log lines (:141-144).  The bar plot is not drawn."""
        if len(all_tokens) == 0:
            return
        k = self.index.ntotal
        lab = torch.as_tensor(all_tokens, dtype=torch.int32, device=self.device)
        counts = torch.empty(k, dtype=torch.int64, device=self.device)
        _lib.check(_lib.load().at_bincount(_lib.ptr(lab), lab.numel(), k, _lib.ptr(counts), _lib.stream_ptr()))
        counts = counts.cpu().numpy()
        self.token_counts = counts
        used = np.nonzero(counts)[0]
        order = used[np.argsort(-counts[used], kind="stable")]
        self.logger.info(f"Total tokens: {len(all_tokens)}")
        self.logger.info(f"Unique tokens: {len(used)}")
        self.logger.info(f"Most common token: [({int(order[0])}, {int(counts[order[0]])})]")
        self.logger.info(f"Least common token: ({int(order[-1])}, {int(counts[order[-1]])})")

    @staticmethod
    def token_statistics(counts):
        """The numbers plot_token_distribution / analyze_zipf_and_tail print (reference :177-236) from a histogram: tokens by
        descending frequency (ties: lower token id first -- the reference keeps first-seen order), the number of tokens that
        account for 80 % of the occurrences, and the least-squares line through the middle 80 % of the log-log rank /
        frequency curve (scipy.stats.linregress in the reference: slope and r^2)."""
        counts = np.asarray(counts, dtype=np.int64)
        used = np.nonzero(counts)[0]
        order = used[np.argsort(-counts[used], kind="stable")]
        freq = counts[order]
        total = int(freq.sum())
        cum = np.cumsum(freq) / total
        tail_start = int(np.searchsorted(cum, 0.8))
        out = dict(tokens=order, frequencies=freq, total=total, top_80_percent=tail_start + 1, tail_start=tail_start,
                   tail_proportion=1 - tail_start / len(freq), slope=float("nan"), r_squared=float("nan"))
        lo, hi = int(0.1 * len(freq)), int(0.9 * len(freq))
        if hi - lo >= 2:
            lx, ly = np.log(np.arange(1, len(freq) + 1))[lo:hi], np.log(freq.astype(np.float64))[lo:hi]
            dx, dy = lx - lx.mean(), ly - ly.mean()
            sxx, sxy, syy = float(dx @ dx), float(dx @ dy), float(dy @ dy)
            out["slope"] = sxy / sxx
            out["r_squared"] = (sxy * sxy) / (sxx * syy) if syy > 0 else 0.0
        return out

    def plot_token_distribution(self, all_tokens):
        """The statistics the reference prints here and in analyze_zipf_and_tail (:177-236), from the device histogram of
        analyze_tokens; the three figures are not drawn."""
        if len(all_tokens) == 0:
            return
        if self.token_counts is None or int(self.token_counts.sum()) != len(all_tokens):
            self.analyze_tokens(all_tokens)
        st = self.token_statistics(self.token_counts)
        tokens, freq = st["tokens"], st["frequencies"]
        print(f"Total unique tokens: {len(tokens)}")
        print(f"Total token occurrences: {st['total']}")
        print(f"Most common token (rank 1): Token {int(tokens[0])} (used {int(freq[0])} times)")
        print(f"Least common token (rank {len(tokens)}): Token {int(tokens[-1])} (used {int(freq[-1])} times)")
        print(f"Top {st['top_80_percent']} tokens account for 80% of all token occurrences")
        print(f"Frequency ratio between most and least common: {freq[0] / freq[-1]:.2f}")
        self.analyze_zipf_and_tail(freq)

    def analyze_zipf_and_tail(self, frequencies):
        """Reference :201-236 on a descending frequency array."""
        counts = np.asarray(frequencies, dtype=np.int64)
        st = self.token_statistics(counts)   # already sorted: the stable ordering keeps it
        print(f"Zipf's law slope: {st['slope']:.2f} (closer to -1 indicates closer fit to Zipf's law)")
        print(f"R-squared value: {st['r_squared']:.2f}")
        print(f"Proportion of tokens in the tail (last 20% of occurrences): {st['tail_proportion']:.2%}")
        print(f"Number of tokens accounting for 80% of occurrences: {st['tail_start']}")


if __name__ == "__main__":
    from audio_tokens_config import AudioTokensConfig   # the reference's own config (resolved from its checkout)

    SpecTokenizer(AudioTokensConfig()).run()
