"""The drop-in stage classes on a small synthetic tree: run() x 3 with the reference's file contract, compared with the
oracle (reference torchaudio calls + FAISS restatement) stage by stage."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SR, L = 22050, 22050 * 3


def _config(tmp_path, **over):
    from stage_config import StageConfig

    cfg = StageConfig()
    cfg.split_file = str(tmp_path / "split.json")
    cfg.dest_spec_path = tmp_path / "spectrograms"
    cfg.source_spec_path = tmp_path / "spectrograms"
    cfg.centroids_path = tmp_path / "output" / "centroids.npy"
    cfg.dest_tokenized_path = str(tmp_path / "tokenized_audio")
    cfg.n_fft, cfg.hop_length, cfg.normalize = 1024, 512, True
    cfg.vocab_size, cfg.niter = 32, 6
    cfg.spectrogram_batch_size = 7   # several batches
    cfg.clustering_batch_size = 10000
    cfg.tokenizer_batch_size = 9
    cfg.sort_files = True            # deterministic file order for the comparison (the reference globs unsorted)
    for k, v in over.items():
        setattr(cfg, k, v)
    (tmp_path / "output").mkdir(exist_ok=True)
    ytids = [f"clip{i:04d}" for i in range(24)]
    json.dump({"train": ytids[:20], "validation": ytids[20:]}, open(cfg.split_file, "w"))
    return cfg, ytids


def _wave(ytid):
    import torch
    from oracle import synth_ref

    idx = int(ytid[4:])
    if idx == 3:
        return torch.zeros(1, L)          # silent clip: NaN after min-max -> dropped like the reference does
    return torch.from_numpy(synth_ref.make_clip(4242, idx, L if idx % 5 else L - 777)).reshape(1, -1)


def _check_tokens_against_the_oracle(wave, clip_ids, tokens, cents, T, what):
    """tokens of the listed clips (a (B*T,) array, T frames per clip) against the CPU oracle end to end: the reference's own
    torchaudio expressions for mel dB + min-max (oracle.mel_ref), normalize_vectors, and the scalar FAISS search formula;
    mismatches only where the oracle's own fp64 top-2 gap is a near tie."""
    from oracle import faiss_ref, mel_ref

    cents = np.ascontiguousarray(cents, dtype=np.float32)
    flips = total = 0
    for b in clip_ids:
        ref_spec = mel_ref.mel_db_torchaudio(wave[b].cpu().numpy(), SR, 1024, 512, 64, True).T.astype(np.float32)
        xn = mel_ref.normalize_rows(np.ascontiguousarray(ref_spec))
        lab, d1, d2 = faiss_ref.assign_l2_scalar(xn, cents)
        _, e1, e2 = faiss_ref.assign_l2_f64(xn, cents)
        got = np.asarray(tokens[b * T:(b + 1) * T])
        mism = got != lab
        # the GPU rows differ from the oracle's by the mel kernel's <= 1e-4 error as well: a flip needs a small gap
        gap = (e2 - e1) / np.maximum(e1, 1e-30)
        assert (gap[mism] < 2e-3).all(), (what, b, float(gap[mism].max()))
        flips += int(mism.sum())
        total += len(got)
    print(f"{what}: {total} tokens of {len(clip_ids)} clips against the oracle, {flips} near-tie flips")
    assert flips <= max(2, total // 200)


def test_three_stages_run_with_the_reference_file_contract(tmp_path, monkeypatch):
    import torch
    from oracle import faiss_ref, mel_ref
    from processors.cluster_creator import ClusterCreator
    from processors.spec_tokenizer import SpecTokenizer
    from processors.spectrogram_generator import SpectrogramGenerator

    cfg, ytids = _config(tmp_path)
    gen = SpectrogramGenerator(cfg)
    monkeypatch.setattr(gen, "find_audio_file", lambda y: None if y == "clip0005" else f"/audio/{y}.flac")
    monkeypatch.setattr(gen, "preprocess_waveform", lambda p: _wave(os.path.basename(p)[:-5]).cuda())
    gen.run()

    # ---- stage 1 files
    train_files = sorted((tmp_path / "spectrograms" / "train").glob("*.npy"))
    names = [f.stem for f in train_files]
    assert "clip0003" not in names and "clip0005" not in names and len(names) == 18   # silent + missing are skipped
    assert len(list((tmp_path / "spectrograms" / "validation").glob("*.npy"))) == 4
    for f in train_files[:6]:
        with open(f, "rb") as fh:
            header = fh.read(128)
        assert b"'fortran_order': True" in header and b"'descr': '<f4'" in header
        s = np.load(f)
        w = _wave(f.stem)[0].numpy()
        assert s.shape == (64, 1 + len(w) // 512) and s.dtype == np.float32
        ref = mel_ref.mel_db_torchaudio(w, SR, 1024, 512, 64, True)
        assert np.abs(s - ref).max() <= 1e-4
    # public helpers keep their meaning
    one = gen.generate_mel_spectrogram(_wave("clip0001").cuda())
    ref_db = mel_ref.mel_db_torchaudio(_wave("clip0001")[0].numpy(), SR, 1024, 512, 64, False)
    assert tuple(one.shape) == ref_db.shape
    assert np.abs(one.cpu().numpy() - ref_db).max() <= 1e-4 * (ref_db.max() - ref_db.min())
    assert gen.check_for_nan_inf(gen.normalize_spectrogram(gen.generate_mel_spectrogram(torch.zeros(1, 4096).cuda())))

    # ---- stage 2
    ClusterCreator(cfg).run()
    cents = np.load(cfg.centroids_path)
    assert cents.shape == (32, 64) and cents.dtype == np.float32 and not np.isfortran(cents)
    np.testing.assert_allclose(np.linalg.norm(cents, axis=1), 1.0, rtol=1e-5)
    x = np.concatenate([np.load(f).T for f in train_files], axis=0).astype(np.float32)
    km = faiss_ref.Kmeans(64, 32, niter=6)
    km.exact_search = True
    km.train(mel_ref.normalize_rows(x))
    ref_c = mel_ref.normalize_rows(km.centroids)
    rel = np.linalg.norm(cents - ref_c, axis=1) / np.linalg.norm(ref_c, axis=1)
    print("centroids within 1e-4 of the oracle's:", float((rel <= 1e-4).mean()))
    assert (rel <= 1e-4).mean() >= 0.9

    # ---- stage 3
    tok = SpecTokenizer(cfg)
    tok.run()
    ix = faiss_ref.IndexFlatL2(64)
    ix.add(cents)
    total = 0
    for split in ("train", "validation"):
        for f in sorted((tmp_path / "spectrograms" / split).glob("*.npy")):
            t = np.load(tmp_path / "tokenized_audio" / split / f"{f.stem}.npy")
            s = np.load(f).T.astype(np.float32)
            assert t.dtype == np.int64 and t.shape == (s.shape[0],)
            xn = mel_ref.normalize_rows(s)
            lab, d1, d2 = faiss_ref.assign_l2_scalar(xn, cents)
            mism = t != lab
            gap = (d2 - d1) / np.maximum(d1, 1e-30)
            assert (gap[mism] < 1e-4).all()
            total += len(t)
    assert tok.token_counts is not None and int(tok.token_counts.sum()) == sum(
        len(np.load(f)) for f in (tmp_path / "tokenized_audio" / "train").glob("*.npy"))
    assert total > 0


def test_hot_path_in_hbm_matches_the_staged_classes(tmp_path):
    """HotPath (no files) == mel -> Kmeans(all rows) -> tokens computed piecewise with the same operators."""
    import torch
    from at_b200 import FlatL2, Kmeans, MelPlan, row_l2norm, synth_clips
    from at_b200.pipeline import HotPath

    B, K = 30, 64
    wave = synth_clips(4242, 0, B, L)
    hp = HotPath(SR, 1024, 512, 64, True, vocab_size=K, niter=5)
    tok, cents, bad = hp.run_device(wave)
    assert int(bad.sum()) == 0 and tok.dtype == torch.int64
    spec, _, l2 = MelPlan(SR, 1024, 512, 64, True).forward(wave, want_l2=True)
    km = Kmeans(64, K, niter=5, max_points_per_centroid=10 ** 9)
    km.train(l2.reshape(-1, 64))
    c2 = row_l2norm(km.centroids_device())
    assert torch.equal(c2, cents)
    ix = FlatL2(64)
    ix.set_centroids(c2)
    t2, _ = ix.search(spec.reshape(-1, 64), l2norm_rows=True, labels_dtype=torch.int64)
    assert torch.equal(t2, tok)
    # host entry point gives the same tokens
    hb = hp.alloc_bufs(B, L, host=True, chunk_clips=8)
    wh = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
    wh.copy_(wave)
    th, ch, bh = hp.run_host(wh, hb, chunk_clips=8)
    assert torch.equal(th.cuda(), tok) and torch.equal(ch.cuda(), cents)
    # ... and against the CPU oracle: tokens of a few clips end to end (reference mel expressions + FAISS search formula on
    # the final centroids), the centroids against the oracle's k-means over the same rows (free-running: share within 1e-4)
    _check_tokens_against_the_oracle(wave, [0, 7, 29], tok.cpu().numpy(), cents.cpu().numpy(), spec.shape[1], "HotPath")
    from oracle import faiss_ref, mel_ref

    kmo = faiss_ref.Kmeans(64, K, niter=5, max_points_per_centroid=10 ** 9)
    kmo.exact_search = True
    kmo.train(l2.reshape(-1, 64).cpu().numpy())
    ref_c = mel_ref.normalize_rows(kmo.centroids)
    rel = np.linalg.norm(cents.cpu().numpy() - ref_c, axis=1) / np.linalg.norm(ref_c, axis=1)
    print("HotPath centroids within 1e-4 of the oracle's:", float((rel <= 1e-4).mean()))
    assert (rel <= 1e-4).mean() >= 0.75   # free-running: one near-tie flip moves two centroids and then spreads


def test_pcm16_host_path_matches_fp32_host_path():
    """16-bit PCM host buffers (half the PCIe bytes, widened on the device as sample / 32768 like torchaudio.load) give
    bit-identical tokens and centroids to the fp32 host path."""
    import torch
    from at_b200 import pcm16_to_f32, synth_clips
    from at_b200.pipeline import HotPath

    B, L, K = 48, 22050, 64
    wave = synth_clips(4242, 0, B, L)
    pcm = (wave * 32768.0).to(torch.int16)
    assert torch.equal(pcm16_to_f32(pcm), wave)
    # the oracle's statement of torchaudio.load's scaling for 16-bit files (sample / 32768)
    assert np.array_equal(pcm16_to_f32(pcm).cpu().numpy(), pcm.cpu().numpy().astype(np.float32) / np.float32(32768.0))
    odd = pcm.reshape(-1)[1:1001].clone()          # unaligned source pointer: scalar path
    assert torch.equal(pcm16_to_f32(odd), wave.reshape(-1)[1:1001])
    hp = HotPath(22050, 1024, 512, 64, True, K, 3)
    outs = []
    for host in (wave.cpu().pin_memory(), pcm.cpu().pin_memory()):
        bufs = hp.alloc_bufs(B, L, host=True, chunk_clips=20, pcm16=host.dtype == torch.int16)
        tok, cen, bad = hp.run_host(host, bufs, chunk_clips=20)
        outs.append((tok.clone(), cen.clone(), bad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_streaming_spectrogram_to_tokens_matches_the_staged_path():
    """C5 shape in miniature: host chunks streamed through mel -> tokenize with fixed centroids (K = 4096) give the
    tokens the staged path (mel of everything, then one search) gives; fp32 and int16 PCM host chunks agree."""
    import torch
    from at_b200 import FlatL2, MelPlan, row_l2norm, synth_clips
    from at_b200.pipeline import HotPath

    B, L, K, CH = 100, 22050, 4096, 24
    wave = synth_clips(4242, 0, B, L)
    plan = MelPlan(22050, 1024, 512, 64, True)
    spec, bad, l2 = plan.forward(wave, want_l2=True)
    rows = l2.reshape(-1, 64)
    cents = row_l2norm(rows[torch.randperm(rows.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(1))[:K]].contiguous())
    ix = FlatL2(64)
    ix.set_centroids(cents)
    want, _ = ix.search(spec.reshape(-1, 64).contiguous(), l2norm_rows=True, want_dist=False, labels_dtype=torch.int64)
    hp = HotPath(22050, 1024, 512, 64, True, K, 1)
    for dtype in (torch.float32, torch.int16):
        host = (wave if dtype == torch.float32 else (wave * 32768.0).to(torch.int16)).cpu().pin_memory()
        chunks = [host[b0:b0 + CH] for b0 in range(0, B, CH)]
        got, flags = [], []
        for tok, bd in hp.stream_tokenize(iter(chunks), cents, chunk_clips=CH):
            got.append(tok.clone()), flags.append(bd.clone())
        got = torch.cat(got)
        assert torch.equal(got, want.cpu())
        assert int(torch.cat(flags).sum()) == 0
        # ... and the streamed tokens against the CPU oracle, end to end, on clips of three different chunks
        _check_tokens_against_the_oracle(wave, [1, 40, 99], got.numpy(), cents.cpu().numpy(), spec.shape[1], f"stream {dtype}")


def test_large_vocab_tensor_search_matches_exact_scan():
    """C4 regime (K = 16384): the tcgen05 search returns the exact fp32 kernel's labels."""
    import torch
    from at_b200 import FlatL2, MelPlan, _lib, row_l2norm, synth_clips

    wave = synth_clips(4242, 0, 400, 22050)
    plan = MelPlan(22050, 1024, 512, 64, True)
    _, _, l2 = plan.forward(wave, want_l2=True)
    rows = l2.reshape(-1, 64).contiguous()          # 400 * 44 = 17,600 rows
    K = 16384
    g = torch.Generator("cuda").manual_seed(7)
    cents = (rows[torch.randperm(rows.shape[0], device="cuda", generator=g)[:K]] +
             0.01 * torch.randn(K, 64, device="cuda", generator=g)).contiguous()
    ix = FlatL2(64)
    ix.set_centroids(cents)
    a, da = ix.search(rows, algo=_lib.ALGO_SIMT)
    b, db = ix.search(rows, algo=_lib.ALGO_TENSOR)
    assert torch.equal(a, b)
    assert torch.equal(da, db)


def test_use_convolution_branch_matches_the_reference_expressions(tmp_path, monkeypatch):
    """config.use_convolution = True (cluster_creator.py:28-34,68-81; spec_tokenizer.py:92-104,115-121): the seeded,
    never-trained Conv1d expansion 64 -> 640 values per frame, k-means and tokenization over the wide rows, against the
    reference's own torch expressions for the expansion and the FAISS restatement for stages 2-3."""
    import torch
    import at_b200
    from oracle import faiss_ref, mel_ref
    from processors.cluster_creator import ClusterCreator
    from processors.spec_tokenizer import SpecTokenizer
    from processors.spectrogram_generator import SpectrogramGenerator

    cfg, ytids = _config(tmp_path, use_convolution=True, num_kernels=10, kernel_size=3, vocab_size=48, niter=5)
    gen = SpectrogramGenerator(cfg)
    monkeypatch.setattr(gen, "find_audio_file", lambda y: f"/audio/{y}.flac")
    monkeypatch.setattr(gen, "preprocess_waveform", lambda p: _wave(os.path.basename(p)[:-5]).cuda())
    gen.run()
    train_files = sorted((tmp_path / "spectrograms" / "train").glob("*.npy"))

    # the reference's layer, built the reference's way (seed, then the default initialisation on the CPU)
    def ref_layer():
        torch.manual_seed(cfg.random_seed)
        return torch.nn.Conv1d(1, cfg.num_kernels, cfg.kernel_size, padding=cfg.kernel_size // 2)

    def ref_apply(conv, rows):   # cluster_creator.py:68-81
        t = torch.tensor(np.array(rows)).float().unsqueeze(1)
        return conv(t).transpose(1, 2).reshape(-1, cfg.num_kernels * cfg.n_mels).detach().numpy()

    cc = ClusterCreator(cfg)
    conv = ref_layer()
    assert torch.equal(cc.conv_weight.cpu(), conv.weight.detach().reshape(10, 3)) and torch.equal(cc.conv_bias.cpu(), conv.bias.detach())
    x = np.concatenate([np.load(f).T for f in train_files], axis=0).astype(np.float32)
    wide_ref = ref_apply(conv, x)
    wide = cc.apply_convolution(x)
    assert tuple(wide.shape) == (x.shape[0], 640)
    assert np.abs(wide.cpu().numpy() - wide_ref).max() <= 2e-6
    # other kernel shapes of the same operator (odd sizes keep the width, like padding = kernel_size // 2)
    for kc, ks in ((4, 5), (1, 1), (16, 7)):
        c2 = torch.nn.Conv1d(1, kc, ks, padding=ks // 2)
        got = at_b200.conv_expand(torch.from_numpy(x[:777]).cuda(), c2.weight.detach().reshape(kc, ks).cuda().contiguous(),
                                  c2.bias.detach().cuda())
        want = c2(torch.from_numpy(x[:777]).unsqueeze(1)).transpose(1, 2).reshape(-1, kc * 64).detach().numpy()
        assert np.abs(got.cpu().numpy() - want).max() <= 5e-6

    # ---- stage 2 over the wide rows
    cc.run()
    cents = np.load(cfg.centroids_path)
    assert cents.shape == (48, 640) and cents.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(cents, axis=1), 1.0, rtol=1e-5)
    km = faiss_ref.Kmeans(640, 48, niter=5)
    km.exact_search = True
    km.train(mel_ref.normalize_rows(wide_ref))
    ref_c = mel_ref.normalize_rows(km.centroids)
    rel = np.linalg.norm(cents - ref_c, axis=1) / np.linalg.norm(ref_c, axis=1)
    print("wide centroids within 1e-4 of the oracle's:", float((rel <= 1e-4).mean()), "max", float(rel.max()))
    assert (rel <= 1e-4).mean() >= 0.9

    # ---- stage 3 over the wide rows
    tok = SpecTokenizer(cfg)
    assert torch.equal(tok.conv_weight, cc.conv_weight) and torch.equal(tok.conv_bias, cc.conv_bias)
    tok.run()
    checked = flips = 0
    for split in ("train", "validation"):
        for f in sorted((tmp_path / "spectrograms" / split).glob("*.npy")):
            t = np.load(tmp_path / "tokenized_audio" / split / f"{f.stem}.npy")
            s = np.load(f).T.astype(np.float32)
            assert t.dtype == np.int64 and t.shape == (s.shape[0],)
            xn = mel_ref.normalize_rows(ref_apply(conv, s))
            lab, d1, d2 = faiss_ref.assign_l2_scalar(xn, cents)
            mism = t != lab
            gap = (d2 - d1) / np.maximum(d1, 1e-30)
            assert (gap[mism] < 1e-4).all()
            checked += len(t)
            flips += int(mism.sum())
    print("wide tokens checked:", checked, "near-tie flips:", flips)
    assert checked > 0 and flips <= checked * 1e-3


def test_wide_row_search_and_update_match_the_oracle():
    """d > 128 (the use_convolution regime): exact tiled fp32 search vs the scalar FAISS restatement, ragged sizes, and one
    teacher-forced Lloyd step at d = 640 (k_gather_sum's wide instantiations)."""
    import torch
    from at_b200 import FlatL2, LloydTrainer, row_l2norm
    from oracle import faiss_ref

    g = torch.Generator("cuda").manual_seed(11)
    for n, d, k in ((1000, 640, 50), (777, 132, 129), (4099, 320, 64), (65, 1024, 3)):
        x = row_l2norm(torch.rand(n, d, device="cuda", generator=g))
        c = x[torch.randperm(n, device="cuda", generator=g)[:k]].contiguous() + 0.01
        ix = FlatL2(d)
        ix.set_centroids(c)
        lab, dist = ix.search(x, labels_dtype=torch.int64)
        ref, d1, d2 = faiss_ref.assign_l2_scalar(x.cpu().numpy(), c.cpu().numpy())
        mism = lab.cpu().numpy() != ref
        gap = (d2 - d1) / np.maximum(d1, 1e-30)
        assert (gap[mism] < 1e-4).all() and mism.mean() <= 1e-2
        np.testing.assert_allclose(dist.cpu().numpy()[~mism], d1[~mism], rtol=1e-3, atol=1e-5)
    n, d, k = 6000, 640, 40
    x = row_l2norm(torch.rand(n, d, device="cuda", generator=g))
    c0 = x[:k].contiguous()
    tr = LloydTrainer(d, k)
    tr.begin(x)
    tr.set_centroids(c0)
    stats = torch.zeros(4, device="cuda")
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    tr.step(x, stats, labels)
    ref = faiss_ref.lloyd_step(x.cpu().numpy(), c0.cpu().numpy(), exact=True)
    got, lab = tr.get_centroids().cpu().numpy(), labels.cpu().numpy()
    mism = lab != ref["labels"]
    touched = np.zeros(k, dtype=bool)
    touched[lab[mism]] = True
    touched[ref["labels"][mism]] = True
    rel = np.linalg.norm(got - ref["centroids"], axis=1) / np.maximum(np.linalg.norm(ref["centroids"], axis=1), 1e-30)
    assert int(stats.cpu().numpy()[1]) == ref["nsplit"] and mism.mean() < 1e-3
    assert (rel[~touched] <= 1e-4).all()


@pytest.mark.parametrize("n,d,k", [(30000, 640, 500), (9000, 128, 1000), (70000, 320, 130), (20000, 1024, 512)])
def test_wide_row_tensor_search_equals_the_exact_wide_row_kernel(n, d, k):
    """d = 64 NS (the use_convolution regime, d = 640): the slice-accumulating tcgen05 search must return the labels of the
    exact fp32 tile kernel (k_assign_gemm), whose sample is checked against the scalar FAISS restatement."""
    import torch
    from at_b200 import FlatL2, _lib, row_l2norm
    from oracle import faiss_ref

    g = torch.Generator(device="cuda").manual_seed(n + d + k)
    base = torch.rand(n, 64, device="cuda", generator=g)
    x = row_l2norm((base.repeat(1, d // 64) * (1.0 + 0.3 * torch.rand(n, d, device="cuda", generator=g))).contiguous())
    c = (x[torch.randperm(n, device="cuda", generator=g)[:k]] + 0.01 * torch.randn(k, d, device="cuda", generator=g)).contiguous()
    ix = FlatL2(d)
    ix.set_centroids(c)
    le, _ = ix.search(x, algo=_lib.ALGO_SIMT, want_dist=False)
    lt, _ = ix.search(x, algo=_lib.ALGO_TENSOR, want_dist=False)
    _, full = ix.tc_stats()
    print(f"wide tensor search n={n} d={d} k={k}: mismatches {int((le != lt).sum())}, rows scanned exactly {full / n:.2%}")
    assert torch.equal(le, lt)
    la, _ = ix.search(x, want_dist=False)   # ALGO_AUTO: whichever it picks, the same labels
    assert torch.equal(la, le)
    idx = torch.linspace(0, n - 1, 1500, device="cuda").long()
    ref, d1, d2 = faiss_ref.assign_l2_scalar(x[idx].cpu().numpy(), c.cpu().numpy())
    mism = lt[idx].cpu().numpy() != ref
    assert ((d2 - d1)[mism] / np.maximum(d1[mism], 1e-30) < 1e-3).all() and mism.mean() < 2e-2


def test_wide_row_lloyd_tensor_vs_exact_bit_identical():
    """k-means over 640-value rows: the tensor path (row image built once, centred on the first centroids' mean) and the exact
    path give the same labels in every iteration, hence bit-identical centroids (exact integer sums)."""
    import torch
    from at_b200 import LloydTrainer, _lib, row_l2norm

    g = torch.Generator(device="cuda").manual_seed(9)
    n, d, k = 60000, 640, 200
    x = row_l2norm(torch.rand(n, 64, device="cuda", generator=g).repeat(1, 10) * (1.0 + 0.3 * torch.rand(n, d, device="cuda", generator=g)))
    init = x[:k].contiguous()
    outs = []
    for algo in (_lib.ALGO_SIMT, _lib.ALGO_TENSOR):
        tr = LloydTrainer(d, k, algo=algo)
        tr.begin(x)
        tr.set_centroids(init)
        lab = torch.empty(n, dtype=torch.int32, device="cuda")
        for _ in range(5):
            tr.step(x, None, lab)
        outs.append((lab.clone(), tr.get_centroids()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
