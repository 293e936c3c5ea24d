#!/usr/bin/env python
"""bench.py -- the audio-tokens hot path on B200: mel spectrogram -> k-means (K=1024, 20 Lloyd iterations over all
frames) -> tokenization, on BASELINE.json's config C2 (20,000 synthetic 10 s clips = 8.62 M frames x 64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the whole hot path over one batch of 20,000 clips per GPU.  Prints ONE JSON line (rank 0).
Multi-GPU: launched by torchrun, one rank per GPU; every rank owns 20,000 clips (weak scaling), k-means runs over
the union of all ranks' frames with one all-reduce of the exact int64 sums per Lloyd iteration.

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
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="B200_PROFILING.md fallback")


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
def cpu_hot_path(n_clips, k, niter, seed=4242, repeat=1):
    """The reference's CPU path on a bounded sample: per-clip torchaudio mel + min-max + NaN check
    (spectrogram_generator.py:63-85), normalize_vectors + Kmeans.train (cluster_creator.py:49-59) and
    IndexFlatL2.search (spec_tokenizer.py:76-78) over the FAISS restatement.  Returns (frames, seconds, stage dict)."""
    import numpy as np
    import torch

    from oracle import faiss_ref, mel_ref, synth_ref

    torch.set_num_threads(os.cpu_count() or 1)
    if torch.cuda.is_available():
        from at_b200 import synth_clips

        wave = synth_clips(seed, 0, n_clips, CLIP_SAMPLES).cpu()
    else:
        wave = torch.from_numpy(synth_ref.make_clips(seed, 0, n_clips, CLIP_SAMPLES))
    mel = mel_ref.TorchaudioMel(SR, N_FFT, HOP, N_MELS, True)
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
    sample = f"{n} clips ({frames} frames) per step: full path incl. {args.niter} Lloyd iterations at K={args.k}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stages_s": stages,
                         "note": "torchaudio calls identical to the reference's; k-means/search = oracle restatement "
                                 "of FAISS 1.8.0 (MKL sgemm via torch), FAISS itself is not installable here"},
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
        "parallelism": f"frames sharded over {world} rank(s); 1 all-reduce of int64 sums/counts per Lloyd iteration",
        "l2_policy": "inputs larger than L2 (17.6 GB waveform, 2.2 GB frames per GPU); no explicit flush",
        "init": "FAISS random-point init (rand_perm(n, 1235)) precomputed on the host outside the timed region",
    }


# ------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime

        os.environ["NCCL_DEBUG"] = os.environ.get("AT_NCCL_DEBUG", "WARN")   # stdout carries exactly one JSON line
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
    hp = HotPath(SR, N_FFT, HOP, N_MELS, True, K, NITER, group=(None if world > 1 else False), algo=args.algo)
    hp.init_rows(n_total)  # host Fisher-Yates, outside the timed region
    wave = synth_clips(4242, rank * B, B, L)
    bufs = hp.alloc_bufs(B, L)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(ev=None):
        if ev:
            ev[0].record()
        spec, bad, l2 = hp.mel(wave, bufs["spec"], bufs["l2"])
        if ev:
            ev[1].record()
        cents = hp.kmeans(l2.reshape(-1, N_MELS), row_offset, n_total)
        from at_b200 import row_l2norm

        cents = row_l2norm(cents)
        if ev:
            ev[2].record()
        hp.tokenize(spec.reshape(-1, N_MELS), cents, bufs["tokens"])
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

    # ---- e2e: the same step from pinned HOST buffers, tokens + centroids read back to the host.  Two host forms of the
    # same clips: fp32 waveforms (what torchaudio.load hands the reference) and the decoder's native 16-bit PCM.
    def agree(ok):
        """True only if every rank says so: whether the e2e leg runs is a collective decision (a rank that skipped it
        alone would leave the others waiting in the leg's all-reduces)."""
        if world == 1:
            return bool(ok)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def measure_e2e(pcm16):
        import psutil

        esz = 2 if pcm16 else 4
        need = B * L * esz
        barrier()   # every rank samples the host memory before any of them allocates
        mem_ok = psutil.virtual_memory().available / max(world, 1) >= 2.5 * need
        if not agree(mem_ok):
            raise MemoryError("not enough host memory for a pinned copy of the waveforms on every rank")
        wave_host, hb, err = None, None, None
        try:
            if pcm16:
                wave_host = torch.empty((B, L), dtype=torch.int16, pin_memory=True)
                for b0 in range(0, B, 2000):   # exact: the synthetic clips are 16-bit-PCM valued
                    wave_host[b0:b0 + 2000].copy_((wave[b0:b0 + 2000] * 32768.0).to(torch.int16))
            else:
                wave_host = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
                wave_host.copy_(wave)
            hb = hp.alloc_bufs(B, L, host=True, pcm16=pcm16)
        except Exception as ex:
            err = ex
        if not agree(err is None):
            del wave_host, hb
            try:
                torch._C._host_emptyCache()
            except Exception:
                pass
            raise RuntimeError(f"host / device buffers for the e2e leg could not be allocated on every rank ({err!r})")
        hb["spec"], hb["l2"], hb["tokens"] = bufs["spec"], bufs["l2"], bufs["tokens"]
        hp.run_host(wave_host, hb, row_offset=row_offset, n_total=n_total)  # warm-up
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        h0.record()
        for _ in range(n_e2e):
            tok_h, cen_h, bad_h = hp.run_host(wave_host, hb, row_offset=row_offset, n_total=n_total)
        h1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n_e2e * 1e3
        dev_ms = h0.elapsed_time(h1) / n_e2e
        tt = torch.tensor([max(wall, dev_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        res = {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(B * L * esz),
               "d2h_bytes_per_step": int(tok_h.numel() * 8 + cen_h.numel() * 4 + bad_h.numel() * 4),
               "api": "at_b200.pipeline.HotPath.run_host (pinned host "
                      + ("int16 PCM" if pcm16 else "fp32 waveforms") + " in, int64 tokens + centroids out)"}
        del wave_host, hb
        try:
            torch._C._host_emptyCache()
        except Exception:
            pass
        return res

    e2e = None
    e2e_pcm16 = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e(False)
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}
        try:
            e2e_pcm16 = measure_e2e(True)
        except Exception as ex:
            e2e_pcm16 = {"value": None, "unit": UNIT, "error": repr(ex)[:200]}

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
                            "one K step carrying the norms) and is bound by the alu pipe of the accumulator scan, not "
                            "by the tensor pipe (profiles/); avg_ms covers every launch of a search (row image when "
                            "rebuilt + scan + candidate re-check + exact scan of the rest); peak = bf16 sustained, "
                            + pk["source"]}
        n_mel, ms_mel = prof["mel"]
        n_upd, ms_upd = prof["update"]
        mel_bytes = B * (L * 4 + T * N_MELS * 4)
        upd_bytes = frames_local * N_MELS * 4 + frames_local * 4 + 2 * K * N_MELS * 4
        rs = {}
        if n_mel:
            g = mel_bytes / (ms_mel / n_mel * 1e-3) / 1e9
            rs["mel"] = {"bound": "hbm", "achieved": g, "peak": pk["hbm"], "unit": "GB/s", "frac": g / pk["hbm"],
                         "avg_ms": ms_mel / n_mel, "frames_per_s": frames_local / (ms_mel / n_mel * 1e-3)}
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
            "roofline": roofline, "roofline_stages": rs, "clocks": clocks, "e2e": e2e, "e2e_pcm16": e2e_pcm16, "gpu_launches": int(launches),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                fr, sec, st = cpu_hot_path(args.cpu_clips, K, NITER)
                line["cpu_baseline"] = {
                    "value": fr / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"{args.cpu_clips} clips ({fr} frames): full path incl. {NITER} Lloyd iterations at K={K}",
                    "stages_s": st,
                    "note": "torchaudio per clip as the reference does; k-means/search = FAISS 1.8.0 restatement (oracle)"}
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
