"""TEST FIXTURE: builds, in a temporary directory, a stand-in for a checkout of the reference with the same import
surface as danavery/audio-tokens' ``run_pipeline.py`` (run_pipeline.py:1-18): a top-level ``audio_tokens_config`` module,
a ``processors`` package holding the three hot-path stage modules, ``model_trainer`` and the rest, and a ``run_pipeline.py``
that imports the four stage classes by the reference's names and runs them in the reference's order.

The stand-in's own three stage modules are decoys that raise on import: a test that gets through ``run_pipeline.py`` has
proved that ``processors.{spectrogram_generator,cluster_creator,spec_tokenizer}`` resolved to audio-tokens_b200/processors/
while ``audio_tokens_config`` and ``processors.model_trainer`` resolved to the checkout.  /root/reference does not exist on
the GPU box and reference sources are never copied into this repository, hence a generated stand-in.
"""
import json
import os
import textwrap

DECOY = 'raise ImportError("the checkout\'s own {name} was imported: the drop-in did not route it")\n'


def make_checkout(root, ytids, n_val=4, **over):
    """Writes the stand-in under ``root`` and returns (checkout_dir, config dict of the paths the stages use)."""
    ck = os.path.join(root, "audio-tokens")
    os.makedirs(os.path.join(ck, "processors"))
    os.makedirs(os.path.join(ck, "output"))
    audio = os.path.join(root, "audioset")
    for y in ytids:
        d = os.path.join(audio, "bal_train", y[:2])
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, f"{y}.flac"), "wb").close()   # torchaudio.load is stubbed by the test; the file must exist
    split = os.path.join(ck, "output", "bal_train_data_split.json")
    json.dump({"train": ytids[:-n_val], "validation": ytids[-n_val:]}, open(split, "w"))
    fields = dict(
        random_seed=4242, split_file=split, audio_source_path=audio, audio_source_sets=["bal_train"],
        dest_spec_path=os.path.join(ck, "spectrograms"), common_sr=22050, normalize=True, n_mels=64, n_fft=1024,
        hop_length=512, spectrogram_batch_size=7, vocab_size=32, niter=6, use_convolution=False, num_kernels=10,
        kernel_size=3, clustering_batch_size=10000, centroids_path=os.path.join(ck, "output", "centroids.npy"),
        source_spec_path=os.path.join(ck, "spectrograms"), dest_tokenized_path=os.path.join(ck, "tokenized_audio"),
        tokenizer_batch_size=9, sort_files=True,
        # fields only the rest of the reference reads: present so that the checkout's config is the one in use
        csv_index_files=[], epochs=1, model_type="lstm", marker="checkout-config")
    fields.update(over)
    with open(os.path.join(ck, "audio_tokens_config.py"), "w") as f:
        f.write("from pathlib import Path\n\n\nclass AudioTokensConfig:\n    def __init__(self):\n")
        for k, v in fields.items():
            if k in ("dest_spec_path", "centroids_path", "source_spec_path"):
                f.write(f"        self.{k} = Path({v!r})\n")
            else:
                f.write(f"        self.{k} = {v!r}\n")
    open(os.path.join(ck, "processors", "__init__.py"), "w").close()
    for name in ("spectrogram_generator", "cluster_creator", "spec_tokenizer"):
        with open(os.path.join(ck, "processors", f"{name}.py"), "w") as f:
            f.write(DECOY.format(name=f"processors/{name}.py"))
    with open(os.path.join(ck, "processors", "model_trainer.py"), "w") as f:
        f.write(textwrap.dedent('''\
            import json
            import os

            from audio_tokens_config import AudioTokensConfig  # noqa: F401  (resolved from the checkout)


            class ModelTrainer:
                """Stand-in for the consumer of the token files: records what it found."""

                def __init__(self, config):
                    self.config = config

                def run(self):
                    out = {"marker": self.config.marker, "config_file": __import__("audio_tokens_config").__file__,
                           "tokens": sorted(os.listdir(os.path.join(self.config.dest_tokenized_path, "train")))}
                    json.dump(out, open(os.path.join(os.path.dirname(str(self.config.centroids_path)), "trainer.json"), "w"))
            '''))
    with open(os.path.join(ck, "run_pipeline.py"), "w") as f:
        f.write(textwrap.dedent('''\
            from audio_tokens_config import AudioTokensConfig
            from processors.cluster_creator import ClusterCreator
            from processors.model_trainer import ModelTrainer
            from processors.spec_tokenizer import SpecTokenizer
            from processors.spectrogram_generator import SpectrogramGenerator


            def main():
                config = AudioTokensConfig()

                SpectrogramGenerator(config).run()
                ClusterCreator(config).run()
                SpecTokenizer(config).run()
                ModelTrainer(config).run()


            if __name__ == "__main__":
                main()
            '''))
    return ck, fields
