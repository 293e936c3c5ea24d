"""The whole hot path on device-resident or host-resident waveforms:

    waveforms -> MelPlan (mel dB + min-max + row-L2 copy) -> k-means (faiss::Clustering semantics, all rows)
              -> nearest-centroid tokens

This is what SpectrogramGenerator.run -> ClusterCreator.run -> SpecTokenizer.run do through .npy files in the
reference (run_pipeline.py:11-13); here the intermediate stays in HBM.  Used by bench.py and by the drop-in
processors when they are chained in one process.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .index import FlatL2
from .kmeans import LloydTrainer, rand_perm
from .mel import MelPlan


class HotPath:
    def __init__(self, sample_rate=22050, n_fft=1024, hop_length=512, n_mels=64, normalize=True, vocab_size=1024,
                 niter=20, group=False, seed=1234, algo=_lib.ALGO_AUTO, reduce="auto"):
        self.plan = MelPlan(sample_rate, n_fft, hop_length, n_mels, normalize)
        self.k, self.d, self.niter, self.group, self.seed = vocab_size, n_mels, niter, group, seed
        self.trainer = LloydTrainer(n_mels, vocab_size, group=group, algo=algo, reduce=reduce)
        self.index = FlatL2(n_mels)
        self.algo = algo
        self._init_rows = {}
        self._init_dev = {}

    # -- pieces ----------------------------------------------------------------------------------
    def init_rows(self, n_total: int) -> np.ndarray:
        """FAISS's random-point initialisation: rows rand_perm(n, seed + 1)[:k] of the training set. Depends only on
        (n, seed), so it is computed once per size (host std::mt19937 Fisher-Yates) and cached."""
        if n_total not in self._init_rows:
            self._init_rows[n_total] = rand_perm(n_total, self.seed + 1)[: self.k].astype(np.int64)
        return self._init_rows[n_total]

    def mel(self, wave, spec=None, l2=None, absmax=None):
        """absmax: CUDA float32[1] that receives max |element| of the L2-normalised copy (what k-means' begin needs), computed
        by the mel kernel on the way out instead of by a separate pass over the rows."""
        if absmax is None:
            return self.plan.forward(wave, out=spec, out_l2=l2, want_l2=True)
        absmax.zero_()
        self.plan.set_absmax_out(absmax)
        try:
            return self.plan.forward(wave, out=spec, out_l2=l2, want_l2=True)
        finally:
            self.plan.set_absmax_out(None)

    def kmeans(self, l2_rows, row_offset=0, n_total=None, stats=None, absmax=None):
        """Lloyd iterations over this rank's L2-normalised rows (n_local, d). Returns device centroids (k, d).
        absmax: see mel()."""
        import torch
        import torch.distributed as dist

        n_local = l2_rows.shape[0]
        n_total = n_local if n_total is None else n_total
        # which of FAISS's k initial rows live in this shard: device index tensors, built once per (size, shard)
        key = (n_total, row_offset, n_local, str(l2_rows.device))
        if key not in self._init_dev:
            rows = self.init_rows(n_total)
            mine = (rows >= row_offset) & (rows < row_offset + n_local)
            self._init_dev[key] = (torch.from_numpy(np.nonzero(mine)[0]).to(l2_rows.device),
                                   torch.from_numpy(rows[mine] - row_offset).to(l2_rows.device))
        idx, src = self._init_dev[key]
        cent = torch.zeros((self.k, self.d), dtype=torch.float32, device=l2_rows.device)
        cent.index_copy_(0, idx, l2_rows.index_select(0, src))
        if self.group is not False and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(cent, group=self.group or None)
        self.trainer.begin(l2_rows, n_total, absmax=absmax)
        self.trainer.set_centroids(cent)
        for it in range(self.niter):
            self.trainer.step(l2_rows, None if stats is None else stats[it])
        return self.trainer.get_centroids()

    def tokenize(self, spec_rows, centroids, tokens=None, trained_rows: bool = False):
        """Nearest-centroid tokens of the (un-normalised) spectrogram rows.  trained_rows=True: these are the rows the trainer
        has just clustered (their L2-normalised copy), so the fp16 operand image of the Lloyd iterations is re-used instead
        of being rebuilt from the fp32 rows (at_index_search_trained_rows)."""
        import torch

        self.index.set_centroids(centroids)
        if (trained_rows and self.algo != _lib.ALGO_SIMT and self.d == 64 and self.k >= 64
                and getattr(self.trainer, "_rows_key", None) is not None
                and self.trainer._rows_key[1][0] == spec_rows.shape[0]):
            lab, _ = self.index.search_trained_rows(self.trainer, spec_rows, l2norm_rows=True, labels_dtype=torch.int64,
                                                    labels=tokens)
            return lab
        lab, _ = self.index.search(spec_rows, l2norm_rows=True, algo=self.algo, want_dist=False,
                                   labels_dtype=torch.int64, labels=tokens)
        return lab

    # -- whole path ------------------------------------------------------------------------------
    def cluster_and_tokenize(self, spec, l2, row_offset=0, n_total=None, stats=None, tokens=None, absmax=None):
        """spec / l2: (B, T, d) device tensors from mel().  Returns (tokens int64 (B*T,), unit-norm centroids)."""
        from . import row_l2norm

        cents = self.kmeans(l2.reshape(-1, self.d), row_offset, n_total, stats, absmax=absmax)
        cents = row_l2norm(cents)  # ClusterCreator saves normalize_vectors(kmeans.centroids) (cluster_creator.py:58-61)
        tok = self.tokenize(spec.reshape(-1, self.d), cents, tokens, trained_rows=True)   # l2 is spec's normalised copy
        return tok, cents

    def run_device(self, wave, bufs=None, row_offset=0, n_total=None, stats=None):
        """wave (B, L) fp32 CUDA.  Returns (tokens int64 (B*T,), centroids (k, d), bad (B,))."""
        import torch

        bufs = bufs or {}
        absmax = bufs.get("absmax")
        if absmax is None:
            absmax = torch.zeros(1, dtype=torch.float32, device=wave.device)
        spec, bad, l2 = self.mel(wave, bufs.get("spec"), bufs.get("l2"), absmax=absmax)
        tok, cents = self.cluster_and_tokenize(spec, l2, row_offset, n_total, stats, bufs.get("tokens"), absmax=absmax)
        return tok, cents, bad

    def _ingest(self, wave_host, bufs, bad, chunk_clips, mel_stream):
        """H2D copy of one batch in chunks on bufs["copy_stream"], widening (int16 PCM) + mel on ``mel_stream`` into
        bufs["spec"] / bufs["l2"]; chunk i+1 is copied while chunk i is transformed.  Returns an event recorded on
        ``mel_stream`` after the last mel launch."""
        import torch

        pcm16 = wave_host.dtype == torch.int16
        B, L = wave_host.shape
        spec, l2 = bufs["spec"], bufs["l2"]
        stage = bufs["stage"]  # (2, chunk_clips, L) device
        copy = bufs["copy_stream"]
        copy.wait_stream(mel_stream)
        ev_free = [None, None]
        sp = _lib.ctypes_stream(mel_stream)
        absmax = bufs.get("absmax")   # raised by every chunk's mel launch (k-means' begin reads it instead of scanning l2)
        if absmax is not None:
            with torch.cuda.stream(mel_stream):
                absmax.zero_()
            self.plan.set_absmax_out(absmax)
        try:
            for ci, b0 in enumerate(range(0, B, chunk_clips)):
                nb = min(chunk_clips, B - b0)
                sb = ci & 1
                with torch.cuda.stream(copy):
                    if ev_free[sb] is not None:
                        copy.wait_event(ev_free[sb])
                    (bufs["stage16"] if pcm16 else stage)[sb, :nb].copy_(wave_host[b0:b0 + nb], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                mel_stream.wait_event(ev)
                if pcm16:
                    _lib.check(self.plan.lib.at_pcm16_to_f32(_lib.ptr(bufs["stage16"][sb]), nb * L, _lib.ptr(stage[sb]), sp))
                _lib.check(self.plan.lib.at_mel_forward(self.plan.h, _lib.ptr(stage[sb]), None, None, L, nb,
                                                        _lib.ptr(spec[b0:]), _lib.ptr(l2[b0:]), _lib.ptr(bad[b0:]), sp))
                ev_free[sb] = torch.cuda.Event()
                ev_free[sb].record(mel_stream)
        finally:
            if absmax is not None:
                self.plan.set_absmax_out(None)
        done = torch.cuda.Event()
        done.record(mel_stream)
        return done

    def run_host(self, wave_host, bufs, chunk_clips=None, row_offset=0, n_total=None):
        """End to end from HOST memory: wave_host (B, L) pinned tensor, fp32 waveforms or the decoder's int16 PCM (half
        the PCIe bytes; widened on the device with at_pcm16_to_f32, bufs from alloc_bufs(..., pcm16=True)).  Chunks are
        copied H2D on a copy stream while the mel kernel works on the previous chunk; tokens (int64) and centroids are
        copied back into the pinned host tensors bufs["tokens_host"], bufs["centroids_host"].  Blocks until they are
        there."""
        import torch

        B = wave_host.shape[0]
        main = torch.cuda.current_stream()
        bad = torch.zeros(B, dtype=torch.int32, device="cuda")
        if chunk_clips is None:
            chunk_clips = bufs["stage"].shape[1]   # what alloc_bufs sized the staging buffers for
        self._ingest(wave_host, bufs, bad, chunk_clips, main)
        tok, cents = self.cluster_and_tokenize(bufs["spec"], bufs["l2"], row_offset, n_total, None, bufs.get("tokens"),
                                               absmax=bufs.get("absmax"))
        bufs["tokens_host"].copy_(tok, non_blocking=True)
        bufs["centroids_host"].copy_(cents, non_blocking=True)
        bufs["bad_host"].copy_(bad, non_blocking=True)
        main.synchronize()
        return bufs["tokens_host"], bufs["centroids_host"], bufs["bad_host"]

    def run_host_stream(self, batches, bufs_pair, chunk_clips=None, row_offset=0, n_total=None):
        """The same end-to-end step for a SEQUENCE of host batches (each one a full k-means + tokenize job, e.g. one split
        or one day's clips): the H2D copy and the mel transform of batch i+1 run on side streams while batch i goes through
        k-means and tokenization, so in steady state a step costs max(PCIe time, device time) instead of their sum.

        batches: iterable of pinned (B, L) tensors (fp32 or int16 PCM), same shape throughout; bufs_pair: two buffer sets
        from alloc_bufs(B, L, host=True, ...).  Yields (tokens_host, centroids_host, bad_host) per batch, in order; the
        pinned tensors belong to the buffer set and are overwritten two batches later."""
        import torch

        main = torch.cuda.current_stream()
        if not hasattr(self, "_mel_stream"):
            self._mel_stream = torch.cuda.Stream()
        mel_stream = self._mel_stream
        it = iter(batches)
        pend = None    # (bufs, bad, event) of the batch whose ingest is in flight

        def start(batch, slot):
            bufs = bufs_pair[slot]
            bad = torch.zeros(batch.shape[0], dtype=torch.int32, device="cuda")
            # everything queued on the main stream so far is older than this batch: the previous occupant of the slot has
            # been read back (its result event was synchronised before it was yielded), `bad` has been zeroed
            mel_stream.wait_stream(main)
            return bufs, bad, self._ingest(batch, bufs, bad, chunk_clips or bufs["stage"].shape[1], mel_stream)

        slot = 0
        first = next(it, None)
        if first is None:
            return
        pend = start(first, slot)
        while pend is not None:
            bufs, bad, ev = pend
            nxt = next(it, None)
            slot ^= 1
            pend = start(nxt, slot) if nxt is not None else None
            main.wait_event(ev)
            tok, cents = self.cluster_and_tokenize(bufs["spec"], bufs["l2"], row_offset, n_total, None, bufs.get("tokens"),
                                                   absmax=bufs.get("absmax"))
            bufs["tokens_host"].copy_(tok, non_blocking=True)
            bufs["centroids_host"].copy_(cents, non_blocking=True)
            bufs["bad_host"].copy_(bad, non_blocking=True)
            bufs["result_event"] = torch.cuda.Event()
            bufs["result_event"].record(main)
            bufs["result_event"].synchronize()
            yield bufs["tokens_host"], bufs["centroids_host"], bufs["bad_host"]

    def stream_tokenize(self, host_chunks, centroids, chunk_clips=None):
        """Spectrogram -> tokens for a stream of HOST chunks with fixed centroids (the unbalanced-train shape: far more
        clips than fit in HBM, SpectrogramGenerator.run + SpecTokenizer.run without the spectrogram files in between).

        host_chunks: iterable of pinned (nb <= chunk_clips, L) tensors, fp32 waveforms or int16 PCM, same L throughout
        (chunk_clips defaults to the first chunk's clip count; a multiple of MelPlan.work_groups() keeps every mel launch
        balanced).
        centroids: (k, d) fp32 CUDA tensor, unit-norm rows as ClusterCreator saves them.
        Yields (tokens int64 pinned host tensor (nb * T,), bad int32 pinned host tensor (nb,)) per chunk, in order; the
        copy of chunk i+1 and the read-back of chunk i-1 overlap the kernels of chunk i.  The yielded tensors are views of
        two pinned slots: the slot of the chunk just yielded is written again (asynchronously) as soon as the generator is
        advanced, so consume or copy a result BEFORE asking for the next one."""
        import torch

        self.index.set_centroids(centroids)
        main = torch.cuda.current_stream()
        copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
        st = None
        prev = None
        for ci, chunk in enumerate(host_chunks):
            nb, L = chunk.shape
            if chunk_clips is None:
                chunk_clips = nb   # capacity of the staging slots = the first chunk's clip count
            assert nb <= chunk_clips
            pcm16 = chunk.dtype == torch.int16
            T = self.plan.num_frames(L)
            if st is None:
                st = dict(
                    L=L,
                    stage=torch.empty((2, chunk_clips, L), dtype=torch.float32, device="cuda"),
                    stage16=torch.empty((2, chunk_clips, L), dtype=torch.int16, device="cuda") if pcm16 else None,
                    spec=torch.empty((2, chunk_clips, T, self.d), dtype=torch.float32, device="cuda"),
                    tok=torch.empty((2, chunk_clips * T), dtype=torch.int64, device="cuda"),
                    bad=torch.zeros((2, chunk_clips), dtype=torch.int32, device="cuda"),
                    tok_h=torch.empty((2, chunk_clips * T), dtype=torch.int64, pin_memory=True),
                    bad_h=torch.empty((2, chunk_clips), dtype=torch.int32, pin_memory=True),
                    free=[None, None], done=[None, None])
            assert L == st["L"], "stream_tokenize: every chunk must have the same clip length"
            sb = ci & 1
            with torch.cuda.stream(copy_in):
                if st["free"][sb] is not None:
                    copy_in.wait_event(st["free"][sb])
                (st["stage16"] if pcm16 else st["stage"])[sb, :nb].copy_(chunk, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_in)
            main.wait_event(ev)
            if st["done"][sb] is not None:
                main.wait_event(st["done"][sb])   # the read-back of this slot's previous tokens has finished
            if pcm16:
                _lib.check(self.plan.lib.at_pcm16_to_f32(_lib.ptr(st["stage16"][sb]), nb * L, _lib.ptr(st["stage"][sb]),
                                                         _lib.stream_ptr()))
            st["bad"][sb].zero_()
            _lib.check(self.plan.lib.at_mel_forward(self.plan.h, _lib.ptr(st["stage"][sb]), None, None, L, nb,
                                                    _lib.ptr(st["spec"][sb]), None, _lib.ptr(st["bad"][sb]),
                                                    _lib.stream_ptr()))
            st["free"][sb] = torch.cuda.Event()
            st["free"][sb].record(main)
            self.index.search(st["spec"][sb, :nb].reshape(-1, self.d), l2norm_rows=True, algo=self.algo, want_dist=False,
                              labels_dtype=torch.int64, labels=st["tok"][sb, :nb * T])
            ev2 = torch.cuda.Event()
            ev2.record(main)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev2)
                st["tok_h"][sb, :nb * T].copy_(st["tok"][sb, :nb * T], non_blocking=True)
                st["bad_h"][sb, :nb].copy_(st["bad"][sb, :nb], non_blocking=True)
                st["done"][sb] = torch.cuda.Event()
                st["done"][sb].record(copy_out)
            if prev is not None:
                psb, pn = prev
                st["done"][psb].synchronize()
                yield st["tok_h"][psb, :pn * T], st["bad_h"][psb, :pn]
            prev = (sb, nb)
        if prev is not None:
            psb, pn = prev
            st["done"][psb].synchronize()
            T = self.plan.num_frames(st["L"])
            yield st["tok_h"][psb, :pn * T], st["bad_h"][psb, :pn]

    def alloc_bufs(self, B, L, host=False, chunk_clips=None, pcm16=False, device_outputs=True):
        """device_outputs=False leaves out spec / l2 / tokens (the caller plugs in buffers it already owns)."""
        import torch

        T = self.plan.num_frames(L)
        if chunk_clips is None:
            chunk_clips = self.plan.work_groups()   # a balanced mel launch per staged chunk
        bufs = {}
        if device_outputs:
            bufs.update(
                spec=torch.empty((B, T, self.d), dtype=torch.float32, device="cuda"),
                l2=torch.empty((B, T, self.d), dtype=torch.float32, device="cuda"),
                tokens=torch.empty(B * T, dtype=torch.int64, device="cuda"),
                absmax=torch.zeros(1, dtype=torch.float32, device="cuda"),
            )
        if host:
            bufs.update(
                stage=torch.empty((2, chunk_clips, L), dtype=torch.float32, device="cuda"),
                copy_stream=torch.cuda.Stream(),
                tokens_host=torch.empty(B * T, dtype=torch.int64, pin_memory=True),
                centroids_host=torch.empty((self.k, self.d), dtype=torch.float32, pin_memory=True),
                bad_host=torch.empty(B, dtype=torch.int32, pin_memory=True),
            )
            if pcm16:
                bufs["stage16"] = torch.empty((2, chunk_clips, L), dtype=torch.int16, device="cuda")
        return bufs
