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
        same log lines.  The matplotlib plots / Zipf fit of the reference are cosmetic and not reproduced."""
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


if __name__ == "__main__":
    from audio_tokens_config import AudioTokensConfig   # the reference's own config (resolved from its checkout)

    SpecTokenizer(AudioTokensConfig()).run()
