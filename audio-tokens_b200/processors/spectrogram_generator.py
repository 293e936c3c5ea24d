"""Stage 1 drop-in: same class, constructor, methods and .npy outputs as the reference's
processors/spectrogram_generator.py, with the per-clip torchaudio transform replaced by one batched launch of the
fused sm_100a mel kernel (at_mel_forward) per ``spectrogram_batch_size`` clips."""
import json
import logging
from concurrent.futures import ThreadPoolExecutor
import os
import shutil
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm

from at_b200 import MelPlan, ResamplePlan


class SpectrogramGenerator:
    def __init__(self, config):
        self.config = config
        self.logger = logging.getLogger(__name__)
        if not torch.cuda.is_available():
            raise RuntimeError("SpectrogramGenerator (B200 build) needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda")
        c = config
        # fused plan used by populate_specs (normalisation inside the kernel when config.normalize)
        self.plan = MelPlan(c.common_sr, c.n_fft, c.hop_length, c.n_mels, bool(c.normalize))
        # un-normalised plan behind generate_mel_spectrogram (the reference method returns plain dB)
        self._plan_db = self.plan if not c.normalize else MelPlan(c.common_sr, c.n_fft, c.hop_length, c.n_mels, False)
        with open(config.split_file, "r") as f:
            self.data_split = json.load(f)
        self._resamplers = {}     # source rate -> ResamplePlan
        self.writer_threads = 8
        self._last_batch = None   # frame-major device tensor of the last populate_specs call

    # ------------------------------------------------------------------ driver (reference :39-61)
    def run(self):
        for split in ["train", "validation"]:
            self.logger.info(f"Creating {split} spectrograms")
            output_dir = Path(self.config.dest_spec_path) / split
            shutil.rmtree(output_dir, ignore_errors=True)
            output_dir.mkdir(parents=True)
            ytids = self.data_split[split]
            bs = self.config.spectrogram_batch_size
            pending = []
            with ThreadPoolExecutor(max_workers=self.writer_threads) as pool:
                for i in tqdm(range(0, len(ytids), bs), total=len(ytids) // bs, position=0):
                    specs = self.populate_specs(ytids[i: i + bs])
                    if not specs:
                        continue
                    # one device->host copy per batch (the reference pays one per clip), then the clips' .npy files are
                    # written by a small thread pool while the next batch is decoded and transformed
                    host = self._last_batch.to("cpu", non_blocking=False).numpy()
                    for spec in specs:
                        ytid = spec["filename"].replace(".flac", "")
                        a, b = spec["rows"]
                        # (n_mels, T) view of a frame-major tile: np.save writes fortran_order=True like the reference
                        pending.append(pool.submit(np.save, output_dir / f"{ytid}.npy", host[a:b].T))
                for f in pending:
                    f.result()
            self.logger.info(f"{split.capitalize()} spectrograms saved to: {output_dir}")

    # ------------------------------------------------------------------ batch body (reference :63-85)
    def populate_specs(self, source_files):
        names, waves = [], []
        for ytid in source_files:
            audio_file_path = self.find_audio_file(ytid)
            if not audio_file_path:
                continue
            waveform = self.preprocess_waveform(audio_file_path)
            if waveform is None:
                continue
            names.append(audio_file_path)
            waves.append(waveform.reshape(-1))
        return self.specs_from_waveforms(names, waves)

    def specs_from_waveforms(self, names, waves):
        """names[i] (path-like) with mono waveform waves[i] (1-D fp32 tensor at common_sr) -> the reference's list
        of {"filename", "spec"} with NaN/Inf clips dropped."""
        if not waves:
            return []
        out, fo, bad = self.plan.forward_ragged(waves)
        self._last_batch = out
        bad = bad.cpu().tolist()
        specs = []
        for i, name in enumerate(names):
            if bad[i] == 2:
                raise RuntimeError(
                    "Argument #4: Padding size should be less than the corresponding input dimension "
                    f"(clip {name} is shorter than n_fft/2+1 samples)")
            if bad[i]:
                self.logger.debug(f"Bad file: {name}")
                continue
            specs.append({"filename": os.path.basename(str(name)), "spec": out[fo[i]:fo[i + 1]].T,
                          "rows": (int(fo[i]), int(fo[i + 1]))})
        return specs

    # ------------------------------------------------------------------ helpers (reference :87-146)
    def find_audio_file(self, ytid):
        audio_file_path = None
        for source_set in self.config.audio_source_sets:
            audio_file_path = Path(f"{self.config.audio_source_path}/{source_set}/{ytid[:2]}/{ytid}.flac")
            if audio_file_path.exists():
                return audio_file_path
        self.logger.debug(f"Audio file not found: {audio_file_path}")
        return None

    def preprocess_waveform(self, audio_file_path):
        import torchaudio

        try:
            waveform, sr = torchaudio.load(audio_file_path)
        except RuntimeError as e:
            if str(e) == "Failed to decode audio.":
                self.logger.info(f"skipping {audio_file_path}: {e}")
                return None
            raise
        waveform = waveform.to(self.device)
        if sr != self.config.common_sr and waveform.dtype == torch.float32:
            # channel mean + sample-rate conversion in one launch (convert_to_mono + resample below, fused)
            return self._resample_plan(sr).forward(waveform.contiguous())
        waveform = self.convert_to_mono(waveform)
        return self.resample(waveform, sr)

    def _resample_plan(self, sr):
        plan = self._resamplers.get(int(sr))
        if plan is None:
            plan = self._resamplers[int(sr)] = ResamplePlan(int(sr), int(self.config.common_sr))
        return plan

    @staticmethod
    def convert_to_mono(waveform):
        if waveform.shape[0] > 1:
            return torch.mean(waveform, dim=0, keepdim=True)
        return waveform

    def resample(self, waveform, sr):
        """(1, L) waveform at sr -> (1, L') at common_sr: torchaudio.transforms.Resample's arithmetic, one cached filter
        bank per source rate (the reference builds a new Resample module per clip)."""
        if sr != self.config.common_sr:
            w = waveform.to(self.device, torch.float32).reshape(-1, waveform.shape[-1]).contiguous()
            if w.shape[0] != 1:   # batched input: channel by channel, like Resample does
                return torch.cat([self._resample_plan(sr).forward(w[i:i + 1]) for i in range(w.shape[0])], dim=0)
            waveform = self._resample_plan(sr).forward(w)
        return waveform

    def generate_mel_spectrogram(self, audio):
        """audio (1, L) -> (n_mels, T) dB mel spectrogram (un-normalised, like the reference method)."""
        spec, bad = self._plan_db.forward(audio.reshape(1, -1).to(self.device, torch.float32).contiguous())
        if int(bad[0]) == 2:
            raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension")
        return spec[0].T

    @staticmethod
    def normalize_spectrogram(spec):
        return (spec - torch.min(spec)) / (torch.max(spec) - torch.min(spec))

    def check_for_nan_inf(self, data, name="data"):
        if torch.isnan(data).any():
            self.logger.debug(f"Warning: NaN values found in {name}")
            return True
        if torch.isinf(data).any():
            self.logger.debug(f"Warning: Inf values found in {name}")
            return True
        return False


if __name__ == "__main__":
    from audio_tokens_config import AudioTokensConfig   # the reference's own config (resolved from its checkout)

    SpectrogramGenerator(AudioTokensConfig()).run()
