#!/usr/bin/env python
"""bench.py -- the audio-tokens hot path on B200: mel spectrogram -> k-means (K=1024, 20 Lloyd iterations over all
frames) -> tokenization, on BASELINE.json's config C2 (20,000 synthetic 10 s clips = 8.62 M frames x 64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the whole hot path over one batch of 20,000 clips per GPU.  Prints ONE JSON line (rank 0).
Multi-GPU: launched by torchrun, one rank per GPU; every rank owns 20,000 clips (weak scaling: `value`), k-means runs over
the union of all ranks' frames with one exchange of the exact int64 sums per Lloyd iteration (a peer-memory kernel over
NVLink, NCCL as the fallback); the `strong` key carries config C3 as written (the SAME 20,000 clips split over the ranks).
Beside the headline the line carries: `parity` (labels vs the exact kernel on every row and vs the CPU oracle on a sample,
teacher-forced Lloyd step, mel error -- all outside the timed region), `roofline` / `roofline_stages`, `e2e` (pinned host
int16 PCM in, tokens out, pipelined over steps) + `e2e_f32`, `library_baseline` (torchaudio / torch on the same B200),
`configs` (C4: K=16384 tokenize, C5: streamed spectrogram+tokenize at K=4096) and `cpu_baseline`.

--impl reference times the reference's CPU implementation of the same path (torchaudio per clip + the FAISS 1.8.0
restatement under oracle/, because FAISS cannot be installed here) on a bounded sample, on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "audio-tokens_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SR, N_FFT, HOP, N_MELS, CLIP_SAMPLES = 22050, 1024, 512, 64, 220500
METRIC = "mel frames/sec through the hot path (mel spectrogram + k-means K=1024 x 20 Lloyd iters over all frames + tokenize)"
UNIT = "frames/s"
# FP32 SIMT peak measured on this pool (profiles/r01_ubench_f32x2.txt: 252.4 lane-flops/clk/SM x 148 SMs x 1.965 GHz)
FP32_PEAK_TFLOPS = 252.4 * 148 * 1.965e9 / 1e12
MEL_FLOPS_PER_FRAME = 30200.0   # real FFT 25,600 + window 1,024 + |X|^2 1,539 + sparse mel 1,996 + dB


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=20000, help="clips per GPU (C2 = 20000)")
    ap.add_argument("--k", type=int, default=1024)
    ap.add_argument("--niter", type=int, default=20)
    ap.add_argument("--cpu-clips", type=int, default=3000, help="clips in the bounded CPU sample (about 10-30 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 exact SIMT, 2 tcgen05")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C4 / C5 extra keys")
    ap.add_argument("--c5-clips", type=int, default=200000, help="clips streamed per GPU in the C5 leg")
    ap.add_argument("--reduce", default="auto", choices=["auto", "peer", "nccl"], help="exchange step of a Lloyd iteration")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    fp32=FP32_PEAK_TFLOPS, source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, fp32=FP32_PEAK_TFLOPS,
                source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------- CPU arm
_CPU_INPUT = {}


def cpu_input(n_clips, seed=4242):
    """The CPU arm's input clips, built by the oracle's own generator (oracle/synth_ref.c, bit-identical to the device
    generator) so that this arm never loads the product library."""
    import torch

    from oracle import synth_ref

    key = (n_clips, seed)
    if key not in _CPU_INPUT:
        _CPU_INPUT.clear()
        _CPU_INPUT[key] = torch.from_numpy(synth_ref.make_clips_c(seed, 0, n_clips, CLIP_SAMPLES))
    return _CPU_INPUT[key]


def cpu_hot_path(n_clips, k, niter, seed=4242, repeat=1):
    """The reference's CPU path on a bounded sample: per-clip torchaudio mel + min-max + NaN check
    (spectrogram_generator.py:63-85), normalize_vectors + Kmeans.train (cluster_creator.py:49-59) and
    IndexFlatL2.search (spec_tokenizer.py:76-78) over the FAISS restatement (blocked MKL sgemm + fused OpenMP top-1,
    OpenMP compute_centroids).  Like the GPU arm, the two FAISS permutations (a function of n and the seed only) are
    drawn outside the timed region.  Returns (frames, seconds, stage dict)."""
    import numpy as np
    import torch

    from oracle import faiss_ref, mel_ref

    torch.set_num_threads(os.cpu_count() or 1)
    wave = cpu_input(n_clips, seed)
    mel = mel_ref.TorchaudioMel(SR, N_FFT, HOP, N_MELS, True)
    n_rows = n_clips * (1 + CLIP_SAMPLES // HOP)
    perm_cache = {(n_rows, 1235): faiss_ref.rand_perm(n_rows, 1235)}
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        specs = []
        for i in range(n_clips):
            s = mel(wave[i])
            if mel.is_bad(s):
                continue
            specs.append(s.numpy())
        t1 = time.perf_counter()
        x = np.concatenate([s.T for s in specs], axis=0).astype(np.float32)
        xn = mel_ref.normalize_rows(x)
        # the benchmark's k-means runs over ALL frames: FAISS's subsampling (max_points_per_centroid = 256) is lifted on
        # both arms, everything else is FAISS's default
        km = faiss_ref.Kmeans(N_MELS, k, niter=niter, verbose=False, gpu=False, max_points_per_centroid=1 << 30)
        km.perm_cache = perm_cache
        km.train(xn)
        cents = mel_ref.normalize_rows(km.centroids)
        t2 = time.perf_counter()
        ix = faiss_ref.IndexFlatL2(N_MELS)
        ix.add(cents)
        _, tok = ix.search(mel_ref.normalize_rows(x), 1)
        t3 = time.perf_counter()
        res = (x.shape[0], t3 - t0, dict(mel_s=t1 - t0, kmeans_s=t2 - t1, tokenize_s=t3 - t2))
        if best is None or res[1] < best[1]:
            best = res
    return best


def cpu_sample_text(n, frames, niter, k):
    return (f"{n} of the workload's 20,000 clips ({frames} frames) per step: full path incl. {niter} Lloyd iterations at "
            f"K={k} over the sample's frames (per-frame cost of mel, of a Lloyd iteration and of tokenization is linear "
            f"in the number of frames, so frames/s carries over to the full size)")


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.cpu_clips
    for _ in range(args.warmup):
        cpu_hot_path(max(8, n // 8), args.k if n // 8 * 431 >= args.k else 64, 1)
    times = []
    frames = 0
    stages = None
    for _ in range(args.steps):
        frames, sec, stages = cpu_hot_path(n, args.k, args.niter)
        times.append(sec)
    sec = sum(times) / len(times)
    value = frames / sec
    cores = os.cpu_count() or 1
    sample = cpu_sample_text(n, frames, args.niter, args.k)
    cfg = workload_config(args, 1)
    cfg["workload"] += f" ; THIS ARM (reference CPU path) times a bounded sample: {sample}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stages_s": stages,
                         "note": "torchaudio calls identical to the reference's; k-means/search = oracle restatement "
                                 "of FAISS 1.8.0 (MKL sgemm via torch + fused OpenMP top-1, OpenMP compute_centroids), "
                                 "FAISS itself is not installable here; input clips from the oracle's own generator "
                                 "(no product code is loaded by this arm)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    T = 1 + CLIP_SAMPLES // HOP
    return {
        "workload": f"C2: {args.clips} synthetic 10 s clips @22050 Hz per GPU -> {args.clips * T} frames x {N_MELS} mel "
                    f"(n_fft={N_FFT}, hop={HOP}, min-max) ; k-means K={args.k}, {args.niter} Lloyd iterations over all "
                    f"frames of all GPUs (no FAISS subsampling) ; tokenize every frame",
        "clips_per_gpu": args.clips, "frames_per_gpu": args.clips * T, "vocab_size": args.k, "niter": args.niter,
        "parallelism": f"frames sharded over {world} rank(s); 1 exchange of the exact int64 sums/counts per Lloyd iteration "
                       f"(peer-memory kernel over NVLink, NCCL all-reduce as the fallback)",
        "l2_policy": "inputs larger than L2 (17.6 GB waveform, 2.2 GB frames per GPU); no explicit flush",
        "init": "FAISS random-point init (rand_perm(n, 1235)) drawn on the host outside the timed region on BOTH arms",
    }



# ------------------------------------------------------------------------------------------------- parity (outside the timed region)
def parity_block(hp, bufs, wave, args, world, cents, oracle_ok):
    """What the timed step produced, checked: (1) the tensor search's tokens against the exact fp32 kernel on EVERY local
    frame (mismatches are only tolerated inside north_star's 1e-6 near-tie carve-out, evaluated in fp64); (2) the same tokens
    against the CPU oracle's scalar FAISS formula on an evenly spread sample; (3) one teacher-forced Lloyd step at the
    benchmark's K against oracle.faiss_ref.lloyd_step; (4) mel frames against the reference's torchaudio calls on the CPU.
    (2)-(4) need the host cores and run on a single-rank job only."""
    import numpy as np
    import torch

    from at_b200 import FlatL2, LloydTrainer, _lib, row_l2norm

    out = {}
    spec_rows = bufs["spec"].reshape(-1, N_MELS)
    n = spec_rows.shape[0]
    tok = bufs["tokens"]   # written by the last timed step with `cents` (the step's unit-norm centroids)
    ix = FlatL2(N_MELS)
    ix.set_centroids(cents)
    lab_ex, _ = ix.search(spec_rows, l2norm_rows=True, algo=_lib.ALGO_SIMT, want_dist=False, labels_dtype=torch.int64)
    mism = torch.nonzero(tok != lab_ex).flatten()
    worst = 0.0
    if mism.numel():
        xm = row_l2norm(spec_rows[mism].contiguous()).double()
        cd = cents.double()
        d = (xm * xm).sum(1, keepdim=True) + (cd * cd).sum(1)[None, :] - 2.0 * xm @ cd.T
        v, _ = torch.topk(d, 2, dim=1, largest=False)
        worst = float(((v[:, 1] - v[:, 0]) / v[:, 1].clamp_min(1e-30)).abs().max())
    out["tokens_vs_exact_fp32_kernel"] = {"rows_checked": int(n), "mismatches": int(mism.numel()),
                                          "largest_fp64_top2_gap_among_mismatches": worst,
                                          "mismatches_outside_1e-6_carve_out": int(mism.numel()) if worst >= 1e-6 else 0}
    if not oracle_ok:
        out["oracle"] = "skipped on a multi-rank job (host cores are shared by the ranks); see the 1-GPU line and tests/test_gpu_fullsize.py"
        return out
    from oracle import faiss_ref, mel_ref

    # (2) CPU oracle on a sample
    m = min(n, 200000)
    idx = torch.linspace(0, n - 1, m, device="cuda").long()
    xs = mel_ref.normalize_rows(spec_rows[idx].cpu().numpy())
    cs = cents.cpu().numpy()
    ref, d1, d2 = faiss_ref.assign_l2_scalar(xs, cs)
    _, e1, e2 = faiss_ref.assign_l2_f64(xs, cs)
    got = tok[idx].cpu().numpy()
    bad = got != ref
    gap32 = (d2 - d1) / np.maximum(d1, 1e-30)
    gap64 = (e2 - e1) / np.maximum(e1, 1e-30)
    outside = bad & (gap32 >= 1e-6)
    out["tokens_vs_oracle"] = {"rows_checked": int(m), "mismatches": int(bad.sum()),
                               "mismatches_outside_1e-6_carve_out": int(outside.sum()),
                               "rate_outside_carve_out": float(outside.mean()),
                               "largest_fp64_top2_gap_among_mismatches": float(gap64[bad].max()) if bad.any() else 0.0,
                               "all_mismatches_are_fp64_near_ties_below_1e-4": bool((gap64[bad] < 1e-4).all())}
    # (3) teacher-forced Lloyd step, K of the benchmark, 1,000 clips' frames
    nt = min(n, 1000 * (1 + CLIP_SAMPLES // HOP))
    x = bufs["l2"].reshape(-1, N_MELS)[:nt].contiguous()
    xh = x.cpu().numpy()
    K = args.k
    c0 = xh[faiss_ref.rand_perm(nt, 1235)[:K]]
    refstep = faiss_ref.lloyd_step(xh, c0, exact=False)
    tr = LloydTrainer(N_MELS, K, algo=args.algo)
    tr.begin(x)
    tr.set_centroids(torch.from_numpy(c0).cuda())
    st = torch.zeros(4, device="cuda")
    labels = torch.empty(nt, dtype=torch.int32, device="cuda")
    tr.step(x, st, labels)
    gotc = tr.get_centroids().cpu().numpy()
    lab = labels.cpu().numpy()
    flips = lab != refstep["labels"]
    touched = np.zeros(K, dtype=bool)
    touched[lab[flips]] = True
    touched[refstep["labels"][flips]] = True
    rel = np.linalg.norm(gotc - refstep["centroids"], axis=1) / np.maximum(np.linalg.norm(refstep["centroids"], axis=1), 1e-30)
    sv = st.cpu().numpy()
    out["teacher_forced_lloyd_step"] = {
        "rows": int(nt), "k": int(K), "label_flips_vs_oracle": int(flips.sum()),
        "centroids_touched_by_a_flip": int(touched.sum()),
        "max_rel_err_untouched_centroids": float(rel[~touched].max()) if (~touched).any() else 0.0,
        "max_rel_err_all_centroids": float(rel.max()), "within_1e-4": bool((rel[~touched] <= 1e-4).all()),
        "nsplit_equal": bool(int(sv[1]) == refstep["nsplit"]),
        "objective_rel_diff": float(abs(sv[0] - refstep["obj"]) / max(abs(refstep["obj"]), 1e-30))}
    del tr
    # (4) mel frames of a few clips against the reference's torchaudio calls (CPU)
    errs = []
    for b in (0, 1, wave.shape[0] // 2, wave.shape[0] - 1):
        refm = mel_ref.mel_db_torchaudio(wave[b].cpu().numpy(), SR, N_FFT, HOP, N_MELS, True).T
        errs.append(float(np.abs(bufs["spec"][b].cpu().numpy() - refm).max()))
    out["mel_max_abs_err_after_minmax"] = {"clips_checked": len(errs), "max": max(errs), "gate": 1e-4}
    return out


# ------------------------------------------------------------------------------------------------- library path on the same B200
def library_baseline(wave, l2_rows, k, pk):
    """SURVEY.md's stated bar: the library path that already runs on Blackwell, timed on the same GPU in the same job.
    (a) torchaudio MelSpectrogram + AmplitudeToDB + per-clip min-max batched over clips (cuFFT + cuBLAS); (b) the reference's
    per-clip loop with device=cuda (spectrogram_generator.py:63-85: one clip per call, two host syncs per clip for the
    NaN/Inf checks); (c) one Lloyd iteration as torch.matmul + argmin + index_add_ (cuBLAS fp32, and with TF32 allowed)."""
    import torch
    import torchaudio

    T = 1 + CLIP_SAMPLES // HOP
    out = {}
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=SR, n_mels=N_MELS, n_fft=N_FFT, hop_length=HOP).cuda()
    todb = torchaudio.transforms.AmplitudeToDB().cuda()

    def batched(w):
        s = todb(mel(w))                                   # (B, n_mels, T)
        mn = s.amin(dim=(1, 2), keepdim=True)
        mx = s.amax(dim=(1, 2), keepdim=True)
        return (s - mn) / (mx - mn)

    chunk, nclips = 250, 2000
    for _ in range(2):
        batched(wave[:chunk])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b0 in range(0, nclips, chunk):
        r = batched(wave[b0:b0 + chunk])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["mel_torchaudio_batched"] = {"frames_per_s": nclips * T / (ms * 1e-3), "clips": nclips, "chunk_clips": chunk,
                                     "ms": ms, "what": "torchaudio MelSpectrogram + AmplitudeToDB + min-max, batched on this B200 (cuFFT + cuBLAS)"}
    del r

    def per_clip(w):   # the reference's populate_specs body with device = cuda
        s = todb(mel(w.reshape(1, -1)).squeeze(0))
        s = (s - torch.min(s)) / (torch.max(s) - torch.min(s))
        return bool(torch.isnan(s).any()) or bool(torch.isinf(s).any())

    for i in range(5):
        per_clip(wave[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n1 = 300
    for i in range(n1):
        per_clip(wave[i])
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    out["mel_reference_per_clip_device_cuda"] = {"frames_per_s": n1 * T / sec, "clips": n1, "ms_per_clip": sec / n1 * 1e3,
                                                 "what": "the reference's per-clip loop (spectrogram_generator.py:63-85) with device=cuda on this B200"}
    del mel, todb

    # (c) one Lloyd iteration with library ops
    n = l2_rows.shape[0]
    c = l2_rows[torch.randperm(n, device="cuda")[:k]].contiguous()

    def lloyd_torch():
        cn = (c * c).sum(1)
        sums = torch.zeros_like(c)
        counts = torch.zeros(k, device="cuda")
        blk = 1 << 19
        for i0 in range(0, n, blk):
            xb = l2_rows[i0:i0 + blk]
            d = cn[None, :] - 2.0 * (xb @ c.T)            # |x|^2 does not change the argmin
            lab = d.argmin(1)
            sums.index_add_(0, lab, xb)
            counts.index_add_(0, lab, torch.ones_like(lab, dtype=torch.float32))
        return sums / counts.clamp_min(1)[:, None]

    for name, tf32 in (("fp32", False), ("tf32", True)):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            lloyd_torch()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                lloyd_torch()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[f"lloyd_iteration_torch_{name}"] = {"iters_per_s": 1e3 / ms, "ms": ms, "rows": int(n), "k": int(k),
                                                     "what": f"torch.matmul ({name}) + argmin + index_add_ per 524,288-row block on this B200"}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------- configs C4 / C5
def config_c4(spec_rows, pk):
    """BASELINE.json configs[3]: large-vocabulary tokenization, K = 16384 centroids over the step's frames."""
    import torch

    from at_b200 import FlatL2, _lib, row_l2norm

    K = 16384
    n = spec_rows.shape[0]
    g = torch.Generator("cuda").manual_seed(7)
    cents = row_l2norm(spec_rows[torch.randperm(n, device="cuda", generator=g)[:K]].contiguous())
    ix = FlatL2(N_MELS)
    ix.set_centroids(cents)
    lab = torch.empty(n, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ix.search(spec_rows, l2norm_rows=True, algo=_lib.ALGO_TENSOR, want_dist=False, labels=lab)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        ix.search(spec_rows, l2norm_rows=True, algo=_lib.ALGO_TENSOR, want_dist=False, labels=lab)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * n * K * N_MELS / (ms * 1e-3) / 1e12
    s0 = ix.tc_stats()
    return {"workload": f"C4: tokenize {n} frames against K={K} centroids (row image build + tcgen05 search + tails, int64 tokens)",
            "ms": ms, "tokens_per_s": n / (ms * 1e-3), "achieved_tflops": tf, "frac_of_sustained_tensor_peak": tf / pk["tensor_sustained"],
            "target_tokens_per_s": 0.39e9}


def config_conv(spec_rows, pk, n_rows=862000, K=500):
    """SURVEY.md section 8f-4 (config.use_convolution, reference defaults num_kernels = 10, kernel_size = 3, vocab_size = 500):
    Conv1d expansion 64 -> 640 values per frame, row normalisation, nearest-centroid search of the wide rows on the
    slice-accumulating tcgen05 kernel against the exact fp32 tile kernel (labels must be equal), one Lloyd iteration."""
    import torch

    import at_b200
    from at_b200 import FlatL2, LloydTrainer, _lib

    x = spec_rows[:n_rows].contiguous()
    n = x.shape[0]
    g = torch.Generator().manual_seed(42)
    conv_w = (torch.rand(10, 3, generator=g) * 2 - 1).mul_(3 ** -0.5).cuda()   # nn.Conv1d's default range for fan-in 3
    conv_b = (torch.rand(10, generator=g) * 2 - 1).mul_(3 ** -0.5).cuda()

    def timed(fn, reps=3):
        for _ in range(2):
            out = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_exp, wide = timed(lambda: at_b200.conv_expand(x, conv_w, conv_b))
    ms_norm, wn = timed(lambda: at_b200.row_l2norm(wide))
    del wide
    ix = FlatL2(640)
    gd = torch.Generator("cuda").manual_seed(3)
    ix.set_centroids(wn[torch.randperm(n, device="cuda", generator=gd)[:K]].contiguous())
    ms_exact, (lab_e, _) = timed(lambda: ix.search(wn, want_dist=False, algo=_lib.ALGO_SIMT), reps=2)
    ms_tc, (lab_t, _) = timed(lambda: ix.search(wn, want_dist=False, algo=_lib.ALGO_TENSOR))
    tr = LloydTrainer(640, K)
    tr.begin(wn)
    tr.set_centroids(wn[:K].contiguous())
    ms_it, _ = timed(lambda: tr.step(wn, None))
    fl = 2.0 * n * K * 640
    return {"workload": f"use_convolution: {n} frames -> Conv1d(1, 10, 3) -> 640-value rows, K={K}",
            "conv_expand_ms": ms_exp, "row_l2norm_ms": ms_norm,
            "search_tensor_ms": ms_tc, "search_tensor_tflops": fl / (ms_tc * 1e-3) / 1e12,
            "search_tensor_frac_of_sustained_tensor_peak": fl / (ms_tc * 1e-3) / 1e12 / pk["tensor_sustained"],
            "search_exact_fp32_ms": ms_exact, "labels_equal_to_exact_kernel": bool(torch.equal(lab_e, lab_t)),
            "lloyd_iteration_ms": ms_it,
            "note": "search_tensor_ms includes the per-call build of the rows' fp16 image (k-means builds it once per training set)"}


def config_c5(hp_plan_args, wave, clips_target, world, pk):
    """BASELINE.json configs[4] per GPU: clips streamed from pinned HOST int16 PCM chunks through spectrogram + tokenize at
    K = 4096 (fixed centroids), tokens read back to the host, nothing else kept (HotPath.stream_tokenize).  The host chunks
    are a small set of distinct synthetic chunks visited cyclically (1.76 TB of distinct audio does not exist here)."""
    import torch

    from at_b200 import row_l2norm
    from at_b200.pipeline import HotPath

    K = 4096
    T = 1 + CLIP_SAMPLES // HOP
    hp = HotPath(SR, N_FFT, HOP, N_MELS, True, K, 1)
    CH = hp.plan.work_groups()   # clips per streamed chunk: one per clip stream of a mel launch
    spec, _, l2 = hp.mel(wave[:2000].contiguous())
    rows = l2.reshape(-1, N_MELS)
    g = torch.Generator("cuda").manual_seed(11)
    cents = row_l2norm(rows[torch.randperm(rows.shape[0], device="cuda", generator=g)[:K]].contiguous())
    del spec, l2, rows
    n_distinct = 8
    host = [torch.empty((CH, CLIP_SAMPLES), dtype=torch.int16, pin_memory=True) for _ in range(n_distinct)]
    for i, h in enumerate(host):
        h.copy_((wave[i * CH:(i + 1) * CH] * 32768.0).to(torch.int16))
    n_chunks = max(4, clips_target // CH)

    def chunks(m):
        for i in range(m):
            yield host[i % n_distinct]

    for _ in hp.stream_tokenize(chunks(4), cents, chunk_clips=CH):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ntok = 0
    for tok, bad in hp.stream_tokenize(chunks(n_chunks), cents, chunk_clips=CH):
        ntok += tok.numel()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    clips = n_chunks * CH
    h2d = clips * CLIP_SAMPLES * 2
    # the device side alone (same kernels, chunk already resident) names the binding resource
    stage = (wave[:CH] * 32768.0).to(torch.int16)
    from at_b200 import FlatL2, _lib, pcm16_to_f32

    ix = FlatL2(N_MELS)
    ix.set_centroids(cents)
    f32 = torch.empty((CH, CLIP_SAMPLES), dtype=torch.float32, device="cuda")
    lab = torch.empty(CH * T, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        if it == 1:
            e0.record()
        pcm16_to_f32(stage, f32)
        sp, _ = hp.plan.forward(f32)
        ix.search(sp.reshape(-1, N_MELS), l2norm_rows=True, algo=_lib.ALGO_TENSOR, want_dist=False, labels=lab)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 5
    dev_clips_per_s = CH / (dev_ms * 1e-3)
    res = {"workload": f"C5 (per GPU): {clips} clips = {clips * T} frames streamed from pinned host int16 PCM through mel + tokenize "
                       f"at K={K}, int64 tokens + bad flags read back ({n_distinct} distinct {CH}-clip chunks visited cyclically)",
           "clips_per_s": clips / sec, "frames_per_s": clips * T / sec, "seconds": sec, "h2d_gb_per_s": h2d / sec / 1e9,
           "d2h_bytes": int(ntok * 8), "device_only_clips_per_s": dev_clips_per_s,
           "binds": "PCIe host-to-device" if dev_clips_per_s > 1.15 * clips / sec else "device kernels",
           "projected_2M_clips_8_gpus_s": 2.0e6 / (8 * clips / sec)}
    del host
    try:
        torch._C._host_emptyCache()
    except Exception:
        pass
    return res


# ------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    nccl_log = None
    if world > 1:
        import datetime

        # stdout carries exactly one JSON line, and NCCL writes its INFO lines to stdout: unless the caller has chosen a
        # NCCL_DEBUG setting, INFO goes to a per-rank file and rank 0 copies the communicator summary (nranks, transports) to
        # stderr after the run, so the rank count stays checkable
        if "NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ["NCCL_DEBUG_SUBSYS"] = os.environ.get("NCCL_DEBUG_SUBSYS", "INIT")
            if "NCCL_DEBUG_FILE" not in os.environ:
                nccl_log = os.path.join(tempfile.gettempdir(), f"at_b200_nccl_{os.getpid()}_rank{rank}.log")
                os.environ["NCCL_DEBUG_FILE"] = nccl_log
        # a mismatched collective should fail in minutes, not after the default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    from at_b200 import _lib, synth_clips
    from at_b200.pipeline import HotPath

    lib = _lib.load()
    B, L, K, NITER = args.clips, CLIP_SAMPLES, args.k, args.niter
    T = 1 + L // HOP
    frames_local = B * T
    n_total = frames_local * world
    row_offset = rank * frames_local
    hp = HotPath(SR, N_FFT, HOP, N_MELS, True, K, NITER, group=(None if world > 1 else False), algo=args.algo,
                 reduce=args.reduce)
    hp.init_rows(n_total)  # host Fisher-Yates, outside the timed region
    wave = synth_clips(4242, rank * B, B, L)
    bufs = hp.alloc_bufs(B, L)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step(ev=None):
        if ev:
            ev[0].record()
        # (the mel kernel also leaves max |normalised element| in bufs["absmax"]: k-means' begin reads it instead of scanning)
        spec, bad, l2 = hp.mel(wave, bufs["spec"], bufs["l2"], absmax=bufs["absmax"])
        if ev:
            ev[1].record()
        cents = hp.kmeans(l2.reshape(-1, N_MELS), row_offset, n_total, absmax=bufs["absmax"])
        from at_b200 import row_l2norm

        cents = row_l2norm(cents)
        last["centroids"] = cents
        if ev:
            ev[2].record()
        hp.tokenize(spec.reshape(-1, N_MELS), cents, bufs["tokens"], trained_rows=True)   # the frames that were just clustered
        if ev:
            ev[3].record()
        return bad

    for _ in range(args.warmup):
        bad = step()
    barrier()
    assert int(bad.sum().item()) == 0
    lib.at_profile_enable(1)
    launches0 = lib.at_kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        step(evs[s])
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = lib.at_kernel_launches() - launches0
    lib.at_profile_enable(0)
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    stage_ms = [sum(evs[s][i].elapsed_time(evs[s][i + 1]) for s in range(args.steps)) / args.steps for i in range(3)]

    import ctypes

    prof = {}
    for tag, name in enumerate(["search", "mel", "update", "finalize"]):
        cnt, tot = ctypes.c_int64(), ctypes.c_double()
        _lib.check(lib.at_profile_summary(tag, ctypes.byref(cnt), ctypes.byref(tot)))
        prof[name] = (cnt.value, tot.value)

    # ---- the k-means update as a full regroup of every row (the incremental form only moves the rows whose label changed,
    # so its time is not a streaming pass over the rows): three extra, untimed-for-the-headline iterations
    upd_full = None
    try:
        l2_rows = bufs["l2"].reshape(-1, N_MELS)
        hp.trainer.set_incremental(False)
        lib.at_profile_enable(1)
        for _ in range(3):
            hp.trainer.step(l2_rows, None)
        torch.cuda.synchronize()
        lib.at_profile_enable(0)
        for tag, name in enumerate(["search", "mel", "update", "finalize"]):
            cnt, tot = ctypes.c_int64(), ctypes.c_double()
            _lib.check(lib.at_profile_summary(tag, ctypes.byref(cnt), ctypes.byref(tot)))
            if name == "update" and cnt.value:
                upd_full = tot.value / cnt.value
        hp.trainer.set_incremental(True)
    except Exception:
        upd_full = None

    # ---- parity of what the timed steps produced (outside the timed region)
    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(hp, bufs, wave, args, world, last["centroids"], oracle_ok=(world == 1))
        except Exception as ex:  # report, never fake
            parity = {"error": repr(ex)[:300]}

    # ---- e2e: the same step through the public host API (HotPath.run_host_stream): pinned HOST buffers in, int64 tokens +
    # centroids + bad flags read back to the host every step; the copy + mel of step i+1 overlap k-means / tokenize of step i.
    # Headline form: the decoder's native 16-bit PCM (what a FLAC decoder produces; widened on the device).  e2e_f32: the fp32
    # waveforms torchaudio.load hands the reference (twice the PCIe bytes).
    def agree(ok):
        """True only if every rank says so: whether the e2e leg runs is a collective decision (a rank that skipped it
        alone would leave the others waiting in the leg's collectives)."""
        if world == 1:
            return bool(ok)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def measure_e2e(pcm16):
        import psutil

        from at_b200 import hostmem

        esz = 2 if pcm16 else 4
        need = B * L * esz
        barrier()   # every rank samples the host memory before any of them allocates
        mem_ok = psutil.virtual_memory().available / max(world, 1) >= 2.5 * need
        if not agree(mem_ok):
            raise MemoryError("not enough host memory for a pinned copy of the waveforms on every rank")
        wave_host, hb2, err = None, None, None
        saved_aff = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        numa = hostmem.bind_to_gpu(local_rank) if world > 1 else {"note": "single rank: no binding"}
        try:
            if pcm16:
                wave_host = torch.empty((B, L), dtype=torch.int16, pin_memory=True)
                for b0 in range(0, B, 2000):   # exact: the synthetic clips are 16-bit-PCM valued
                    wave_host[b0:b0 + 2000].copy_((wave[b0:b0 + 2000] * 32768.0).to(torch.int16))
            else:
                wave_host = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
                wave_host.copy_(wave)
            hb1 = hp.alloc_bufs(B, L, host=True, pcm16=pcm16, device_outputs=False)
            hb1["spec"], hb1["l2"], hb1["tokens"], hb1["absmax"] = bufs["spec"], bufs["l2"], bufs["tokens"], bufs["absmax"]
            hb2 = hp.alloc_bufs(B, L, host=True, pcm16=pcm16)
        except Exception as ex:
            err = ex
        if not agree(err is None):
            del wave_host, hb2
            try:
                torch._C._host_emptyCache()
            except Exception:
                pass
            if saved_aff is not None:
                os.sched_setaffinity(0, saved_aff)
            raise RuntimeError(f"host / device buffers for the e2e leg could not be allocated on every rank ({err!r})")
        pair = [hb1, hb2]
        for _ in hp.run_host_stream([wave_host, wave_host], pair, row_offset=row_offset, n_total=n_total):   # warm-up
            pass
        barrier()
        n_e2e = 4
        t0 = time.perf_counter()
        for tok_h, cen_h, bad_h in hp.run_host_stream([wave_host] * n_e2e, pair, row_offset=row_offset, n_total=n_total):
            pass
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n_e2e * 1e3
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        # the ceiling the host gives this rank while every rank copies at once: the same pinned buffer through plain
        # cudaMemcpyAsync, no kernels (per-root PCIe / host-memory limit of the box)
        barrier()
        raw = torch.empty((256, L), dtype=wave_host.dtype, device="cuda")
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        nrep = max(1, min(B // 256, 24))
        for i in range(nrep):
            raw.copy_(wave_host[i * 256:(i + 1) * 256], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        raw_gbs = nrep * 256 * L * esz / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del raw
        # one step alone (nothing to overlap with), for the latency of a single job
        t0 = time.perf_counter()
        hp.run_host(wave_host, hb1, row_offset=row_offset, n_total=n_total)
        single_ms = (time.perf_counter() - t0) * 1e3
        res = {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": n_e2e,
               "single_step_latency_ms": single_ms,
               "h2d_bytes_per_step": int(B * L * esz),
               "d2h_bytes_per_step": int(tok_h.numel() * 8 + cen_h.numel() * 4 + bad_h.numel() * 4),
               "h2d_gb_per_s_per_rank": B * L * esz / (e2e_ms * 1e-3) / 1e9,
               "raw_h2d_gb_per_s_this_rank_all_ranks_copying": raw_gbs, "host_placement": numa,
               "api": "at_b200.pipeline.HotPath.run_host_stream (pinned host "
                      + ("int16 PCM" if pcm16 else "fp32 waveforms") + " in, int64 tokens + centroids + bad flags out "
                      "every step; copy + mel of step i+1 overlap k-means + tokenize of step i)"}
        del wave_host, hb1, hb2, pair
        try:
            torch._C._host_emptyCache()
        except Exception:
            pass
        if saved_aff is not None:
            os.sched_setaffinity(0, saved_aff)
        return res

    e2e = None
    e2e_f32 = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e(True)
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}
        try:
            e2e_f32 = measure_e2e(False)
        except Exception as ex:
            e2e_f32 = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

    # ---- config C3 as written: the SAME 20,000 clips split over the ranks (strong scaling)
    strong = None
    if world > 1:
        try:
            Bs = B // world
            wave_s = synth_clips(4242, rank * Bs, Bs, L)
            hp_s = HotPath(SR, N_FFT, HOP, N_MELS, True, K, NITER, group=None, algo=args.algo, reduce=args.reduce)
            nts = Bs * T * world
            hp_s.init_rows(nts)
            bs = hp_s.alloc_bufs(Bs, L)

            def step_s(ev=None):
                if ev:
                    ev[0].record()
                spec_s, _, l2_s = hp_s.mel(wave_s, bs["spec"], bs["l2"], absmax=bs["absmax"])
                if ev:
                    ev[1].record()
                hp_s.cluster_and_tokenize(spec_s, l2_s, rank * Bs * T, nts, None, bs["tokens"], absmax=bs["absmax"])
                if ev:
                    ev[2].record()

            for _ in range(3):
                step_s()
            barrier()
            ns = max(args.steps, 3)
            evs_s = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(ns)]
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(ns):
                step_s(evs_s[i])
            s1.record()
            barrier()
            tms = torch.tensor([s0.elapsed_time(s1) / ns], dtype=torch.float64, device="cuda")
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            sms = float(tms.item())
            strong = {"scaling": "strong", "workload": f"C3: the C2 workload ({Bs * world} clips, {nts} frames) split over {world} ranks",
                      "value": nts / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms, "steps": ns,
                      "mel_ms": sum(e[0].elapsed_time(e[1]) for e in evs_s) / ns,
                      "kmeans_tokenize_ms": sum(e[1].elapsed_time(e[2]) for e in evs_s) / ns,
                      "exchange": hp_s.trainer.reduce}
            del hp_s, bs, wave_s
            torch.cuda.empty_cache()
        except Exception as ex:
            strong = {"error": repr(ex)[:300]}

    # ---- library path on the same GPU, C4 / C5 (single-rank job; C5 also per rank on a multi-rank job)
    lib_base, cfgs = None, {}
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            lib_base = library_baseline(wave, bufs["l2"].reshape(-1, N_MELS), K, peaks())
        except Exception as ex:
            lib_base = {"error": repr(ex)[:300]}
    if not args.no_configs:
        if rank == 0 and world == 1:
            try:
                cfgs["C4"] = config_c4(bufs["spec"].reshape(-1, N_MELS), peaks())
            except Exception as ex:
                cfgs["C4"] = {"error": repr(ex)[:300]}
            try:
                cfgs["use_convolution"] = config_conv(bufs["spec"].reshape(-1, N_MELS), peaks())
            except Exception as ex:
                cfgs["use_convolution"] = {"error": repr(ex)[:300]}
        c5_err, c5 = None, None
        try:
            barrier()
            c5 = config_c5(None, wave, args.c5_clips, world, peaks())
        except Exception as ex:
            c5_err = repr(ex)[:300]
        if world > 1:
            # every rank streams its own clips at the same time (shared host memory / PCIe roots): aggregate = sum
            v = torch.tensor([c5["clips_per_s"] if c5 else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(v)
            if c5:
                c5["clips_per_s_all_ranks"] = float(v.item())
                c5["projected_2M_clips_this_job_s"] = 2.0e6 / max(float(v.item()), 1e-9)
        cfgs["C5"] = c5 if c5 else {"error": c5_err}

    if rank == 0:
        pk = peaks()
        n_search, ms_search = prof["search"]
        flops = 2.0 * frames_local * K * N_MELS
        ach = flops / (ms_search / max(n_search, 1) * 1e-3) / 1e12 if n_search else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("search_dram_bytes_per_launch")
        roofline = {"kernel": "k_assign_tc (tcgen05 distance-argmin)" if args.algo != 1 else "k_assign_simt",
                    "bound": "tensor", "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": (ach / pk["tensor_sustained"]) if ach else None, "traffic": traffic,
                    "launches": n_search, "avg_ms": ms_search / max(n_search, 1),
                    "algorithmic_flops_per_launch": flops,
                    "note": "algorithmic 2*N*K*D flops; the kernel executes 1.25x as many fp16 MMA flops (four K steps + "
                            "one K step carrying the norms); the accumulator scan (1.26 min/max per score on the alu pipe, "
                            "81 % busy) and the 1 kW power cap (SM clock ~1.62 GHz inside the kernel) bind, not the tensor "
                            "pipe (64 % busy; profiles/r02_ncu_full_summary.txt); avg_ms covers every launch of a search "
                            "(row image when rebuilt + scan + candidate re-check of 1.4 % of the rows + exact scan of 0.02 "
                            "%); peak = bf16 sustained, " + pk["source"]}
        n_mel, ms_mel = prof["mel"]
        n_upd, ms_upd = prof["update"]
        mel_bytes = B * (L * 4 + T * N_MELS * 4)
        upd_bytes = frames_local * N_MELS * 4 + frames_local * 4 + 2 * K * N_MELS * 4
        rs = {}
        if n_mel:
            g = mel_bytes / (ms_mel / n_mel * 1e-3) / 1e9
            tfl = frames_local * MEL_FLOPS_PER_FRAME / (ms_mel / n_mel * 1e-3) / 1e12
            rs["mel"] = {"bound": "hbm", "achieved": g, "peak": pk["hbm"], "unit": "GB/s", "frac": g / pk["hbm"],
                         "avg_ms": ms_mel / n_mel, "frames_per_s": frames_local / (ms_mel / n_mel * 1e-3),
                         "fp32_bound": {"achieved": tfl, "peak": pk["fp32"], "unit": "TFLOP/s", "frac": tfl / pk["fp32"],
                                        "note": "algorithmic 30.2 kflop/frame (SURVEY 8d) against the measured FP32 FMA peak "
                                                "(profiles/r01_ubench_f32x2.txt); the two bounds are within 15 % of each other"}}
        if n_upd:
            rs["update"] = {"bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None,
                            "avg_ms": ms_upd / n_upd,
                            "note": "average over the step's Lloyd iterations: the first regroups every row, the others "
                                    "move only the rows whose label changed (exact integer sums, bit-identical result), "
                                    "so this is not a streaming pass; update_full_regroup is"}
        if upd_full:
            g = upd_bytes / (upd_full * 1e-3) / 1e9
            rs["update_full_regroup"] = {"bound": "hbm", "achieved": g, "peak": pk["hbm"], "unit": "GB/s",
                                         "frac": g / pk["hbm"], "avg_ms": upd_full,
                                         "algorithmic_bytes": upd_bytes}
        n_fin, ms_fin = prof["finalize"]
        lloyd_ms = (ms_search * (NITER / (NITER + 1.0)) / args.steps + ms_upd / args.steps + ms_fin / args.steps) / NITER
        # the two GEMM-shaped stages against the same tensor bound as the search kernel: one Lloyd iteration (search + update
        # + finalize, stage time / NITER incl. the once-per-training-set work) and the tokenization of every frame
        it_ms = stage_ms[1] / NITER
        rs["lloyd_iteration"] = {"bound": "tensor", "achieved": flops / (it_ms * 1e-3) / 1e12, "peak": pk["tensor_sustained"],
                                 "unit": "TFLOP/s", "frac": flops / (it_ms * 1e-3) / 1e12 / pk["tensor_sustained"],
                                 "avg_ms": it_ms, "note": "2*N*K*D flops of the iteration's search over kmeans_ms / NITER "
                                                          "(row image, first full regroup and set-up amortised in)"}
        rs["tokenize"] = {"bound": "tensor", "achieved": flops / (stage_ms[2] * 1e-3) / 1e12, "peak": pk["tensor_sustained"],
                          "unit": "TFLOP/s", "frac": flops / (stage_ms[2] * 1e-3) / 1e12 / pk["tensor_sustained"],
                          "avg_ms": stage_ms[2], "note": "2*N*K*D flops over the tokenize stage (operand rebuild + search of "
                                                         "the trained rows' image + int64 labels)"}
        line = {
            "metric": METRIC, "value": n_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (search contraction: fp16 x fp16 tcgen05 MMA, fp32 accumulate, every label certified against or re-evaluated with the fp32 formula)",
            "data": "synthetic", "config": workload_config(args, world),
            "stages": {
                "mel_frames_per_s": frames_local * world / (stage_ms[0] * 1e-3),
                "lloyd_iters_per_s": NITER / (stage_ms[1] * 1e-3),
                "tokens_per_s": frames_local * world / (stage_ms[2] * 1e-3),
                "mel_ms": stage_ms[0], "kmeans_ms": stage_ms[1], "tokenize_ms": stage_ms[2],
                "lloyd_iter_kernel_ms": lloyd_ms, "note": "rank-0 stage times; k-means rows = all ranks' frames",
            },
            "roofline": roofline, "roofline_stages": rs, "clocks": clocks, "e2e": e2e, "e2e_f32": e2e_f32,
            "gpu_launches": int(launches), "parity": parity, "library_baseline": lib_base, "configs": cfgs,
            "collective": (f"Lloyd-iteration exchange: {hp.trainer.reduce}" + (" (at_peer_reduce: one signal + wait + sum kernel "
                           "over CUDA-IPC peer windows, NVLink)" if hp.trainer.reduce == "peer" else "")
                           + "; NCCL for set-up collectives (initial centroids, barriers)") if world > 1 else "none (1 rank)",
        }
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            try:
                fr, sec, st = cpu_hot_path(args.cpu_clips, K, NITER)
                line["cpu_baseline"] = {
                    "value": fr / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": cpu_sample_text(args.cpu_clips, fr, NITER, K),
                    "stages_s": st,
                    "note": "torchaudio per clip as the reference does; k-means/search = FAISS 1.8.0 restatement (oracle)"}
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0 and nccl_log and os.path.exists(nccl_log):
            # the communicator summary, for whoever counts ranks: stderr, so stdout stays one JSON line
            keep = [ln.rstrip() for ln in open(nccl_log, errors="replace")
                    if ("nranks" in ln or "Init COMPLETE" in ln or "NVLS" in ln or "via P2P" in ln)]
            for ln in keep[:40]:
                print(ln, file=sys.stderr)
        if nccl_log:
            try:
                os.unlink(nccl_log)
            except OSError:
                pass


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    run_b200(args)


if __name__ == "__main__":
    main()
