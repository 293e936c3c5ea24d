"""Stage 2 drop-in: same class and outputs as the reference's processors/cluster_creator.py; faiss.Kmeans is
replaced by at_b200.Kmeans (faiss::Clustering semantics on sm_100a kernels)."""
import logging
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm

import at_b200
import at_b200.faiss_compat as faiss
from at_b200.npyio import load_spec_batch


def _set_seed(seed=42):
    """utils/set_seed.py of the reference (FAISS itself ignores it: its seeds stay 1234 / 1235)."""
    import random

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


class ClusterCreator:
    def __init__(self, config):
        self.logger = logging.getLogger(__name__)
        self.config = config
        _set_seed(self.config.random_seed)
        self.gpu = faiss.get_num_gpus() > 0
        if not self.gpu:
            raise RuntimeError("ClusterCreator (B200 build) needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda")
        if self.config.use_convolution:
            # same seeded default initialisation as the reference's nn.Conv1d (:28-34); applied by at_conv_expand
            self.conv_weight, self.conv_bias = at_b200.make_conv_layer(self.config)

    def run(self):
        n_freq_bins = self.config.n_mels
        if self.config.use_convolution:
            n_freq_bins *= self.config.num_kernels
        self.logger.info("starting clustering")
        extra = {}
        mppc = getattr(self.config, "max_points_per_centroid", None)
        if mppc is not None:
            extra["max_points_per_centroid"] = int(mppc)
        kmeans = faiss.Kmeans(n_freq_bins, self.config.vocab_size, niter=self.config.niter, verbose=True,
                              gpu=self.gpu, **extra)
        for i, batch in enumerate(self._batch_generator(self.config.clustering_batch_size)):
            batch = torch.from_numpy(batch).to(self.device)
            if self.config.use_convolution:
                batch = self.apply_convolution(batch)
            batch = at_b200.row_l2norm(batch)  # normalize_vectors on the device
            if i == 0:
                kmeans.train(batch)
            else:
                kmeans.train(batch, init_centroids=kmeans.centroids)
        centroids = self.normalize_vectors(kmeans.centroids)
        self.logger.info(f"Centroids shape: {centroids.shape}")
        np.save(self.config.centroids_path, centroids)
        self.visualize_centroids(centroids)

    def normalize_vectors(self, vectors):
        """v / (||v|| + 1e-10) per row (reference :64-66), computed by at_row_l2norm."""
        v = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(self.device)
        return at_b200.row_l2norm(v).cpu().numpy()

    def apply_convolution(self, time_slice_batch):
        """(n, n_mels) -> (n, n_mels * num_kernels) like the reference (:68-81); CUDA tensor in, CUDA tensor out (numpy is
        uploaded first)."""
        if isinstance(time_slice_batch, np.ndarray) or not torch.is_tensor(time_slice_batch):
            time_slice_batch = torch.from_numpy(np.ascontiguousarray(time_slice_batch, dtype=np.float32)).to(self.device)
        return at_b200.conv_expand(time_slice_batch.contiguous(), self.conv_weight, self.conv_bias)

    def _files(self):
        spec_dir = Path(self.config.source_spec_path) / "train"
        files = list(spec_dir.glob("*.npy"))
        return sorted(files) if getattr(self.config, "sort_files", False) else files

    def _batch_generator(self, batch_size):
        files = self._files()
        for i in tqdm(range(0, len(files), batch_size)):
            # np.concatenate([np.load(f).T ...]).astype(float32) of the reference, with the reads on a thread pool
            yield load_spec_batch(files[i:i + batch_size])[0]

    def visualize_centroids(self, centroids):
        """PCA scatter plot (cosmetic): produced only when matplotlib and scikit-learn are importable."""
        try:
            import matplotlib.pyplot as plt
            from sklearn.decomposition import PCA
        except ImportError:
            self.logger.info("Centroids visualization skipped (matplotlib / scikit-learn not installed)")
            return
        centroids_2d = PCA(n_components=2).fit_transform(centroids)
        plt.figure(figsize=(10, 8))
        plt.scatter(centroids_2d[:, 0], centroids_2d[:, 1])
        plt.title("2D PCA of Centroids")
        plt.savefig("output/centroids_visualization.png")
        plt.close()
        self.logger.info("Centroids visualization saved")


if __name__ == "__main__":
    from audio_tokens_config import AudioTokensConfig   # the reference's own config (resolved from its checkout)

    ClusterCreator(AudioTokensConfig()).run()
