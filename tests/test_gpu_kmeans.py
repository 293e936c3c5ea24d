"""GPU parity of stages 2-3 (at_index_*, at_kmeans_* through the C ABI) against the FAISS restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(n_clips=40, L=22050 * 2, seed=4242):
    """L2-normalised mel frames from synthetic clips (the real input distribution of stages 2-3)."""
    import torch
    from at_b200 import MelPlan, synth_clips

    plan = MelPlan(22050, 1024, 512, 64, True)
    w = synth_clips(seed, 0, n_clips, L)
    spec, bad, l2 = plan.forward(w, want_l2=True)
    assert (bad == 0).all()
    return spec.reshape(-1, 64).contiguous(), l2.reshape(-1, 64).contiguous()


def _check_labels(labels, x, c, what):
    """Token parity gate: equal to the fp32 FAISS-formula oracle except rows whose top-2 gap is tiny."""
    from oracle import faiss_ref

    ref, d1, d2 = faiss_ref.assign_l2_scalar(x, c)
    l64, e1, e2 = faiss_ref.assign_l2_f64(x, c)
    gap32 = (d2 - d1) / np.maximum(d1, 1e-30)
    gap64 = (e2 - e1) / np.maximum(e1, 1e-30)
    mism = labels != ref
    rate_1e6 = float((mism & (gap32 >= 1e-6)).mean())
    print(f"{what}: n={len(labels)} mismatches={int(mism.sum())} outside-1e-6-carve-out rate={rate_1e6:.2e}")
    # every mismatching row must be a near tie of the expanded fp32 formula
    assert (gap64[mism] < 1e-4).all(), (what, gap64[mism].max())
    assert rate_1e6 < 1e-4
    return ref


def test_row_l2norm_matches_numpy():
    import torch
    from at_b200 import row_l2norm
    from oracle import mel_ref

    x = torch.rand(1000, 64, device="cuda")
    x[3] = 0
    got = row_l2norm(x).cpu().numpy()
    np.testing.assert_allclose(got, mel_ref.normalize_rows(x.cpu().numpy()), rtol=2e-6, atol=1e-9)
    assert (got[3] == 0).all()
    for d in (7, 16, 100):
        y = torch.randn(33, d, device="cuda")
        np.testing.assert_allclose(row_l2norm(y).cpu().numpy(), mel_ref.normalize_rows(y.cpu().numpy()), rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("k", [1, 5, 256, 500])
def test_search_simt_matches_oracle(k):
    import torch
    from at_b200 import FlatL2, _lib

    spec, l2 = _frames(20)
    x = l2[:20000]
    c = x[torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(k))[:k].cuda()].contiguous()
    ix = FlatL2(64)
    ix.set_centroids(c)
    lab, dist = ix.search(x, algo=_lib.ALGO_SIMT)
    ref = _check_labels(lab.cpu().numpy(), x.cpu().numpy(), c.cpu().numpy(), f"simt k={k}")
    # fused normalisation == normalise then search
    lab2, dist2 = ix.search(spec[:20000], l2norm_rows=True, algo=_lib.ALGO_SIMT)
    assert torch.equal(lab, lab2) and torch.equal(dist, dist2)
    # int64 labels
    lab3, _ = ix.search(x, algo=_lib.ALGO_SIMT, labels_dtype=torch.int64)
    assert lab3.dtype == torch.int64 and torch.equal(lab3.int(), lab)
    from oracle import faiss_ref

    _, d1, _ = faiss_ref.assign_l2_f64(x.cpu().numpy(), c.cpu().numpy())
    np.testing.assert_allclose(dist.cpu().numpy(), d1, rtol=5e-3, atol=5e-7)


@pytest.mark.parametrize("d", [8, 33, 128])
def test_search_other_dims(d):
    import torch
    from at_b200 import FlatL2, _lib

    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.rand(3001, d, device="cuda", generator=g)
    c = torch.rand(77, d, device="cuda", generator=g)
    ix = FlatL2(d)
    ix.set_centroids(c)
    lab, dist = ix.search(x, algo=_lib.ALGO_SIMT)
    _check_labels(lab.cpu().numpy(), x.cpu().numpy(), c.cpu().numpy(), f"simt d={d}")


def test_lowest_index_wins_exact_ties_and_empty_input():
    import torch
    from at_b200 import FlatL2

    c = torch.zeros(4, 64, device="cuda")
    c[1] = 1.0
    c[3] = 1.0
    ix = FlatL2(64)
    ix.set_centroids(c)
    lab, _ = ix.search(torch.ones(130, 64, device="cuda"))
    assert (lab == 1).all()
    lab, _ = ix.search(torch.ones(0, 64, device="cuda"))
    assert lab.numel() == 0


def test_faiss_like_index_api():
    import torch
    from at_b200 import IndexFlatL2
    from oracle import faiss_ref

    _, l2 = _frames(60)
    x = l2[:5000].cpu().numpy()
    c = x[::50].copy()
    a = IndexFlatL2(64)
    a.add(c)
    D, I = a.search(x, 1)
    assert D.shape == (5000, 1) and I.shape == (5000, 1) and I.dtype == np.int64 and D.dtype == np.float32
    assert a.ntotal == len(c)
    _check_labels(I[:, 0], x, c, "IndexFlatL2")


def test_lloyd_single_step_teacher_forced():
    """Tier (i) of the centroid gate: from the oracle's centroids C_t, one device step gives C_{t+1} within
    1e-4 (relative L2 per centroid) on every centroid not touched by a label mismatch; same nsplit."""
    import torch
    from at_b200 import LloydTrainer
    from oracle import faiss_ref

    _, l2 = _frames(40)
    x = l2.cpu().numpy()
    n, k = x.shape[0], 256
    cents = x[faiss_ref.rand_perm(n, 1235)[:k]]
    tr = LloydTrainer(64, k)
    tr.begin(l2)
    stats = torch.zeros(4, device="cuda")
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    for it in range(4):
        ref = faiss_ref.lloyd_step(x, cents, exact=True)
        tr.set_centroids(torch.from_numpy(cents).cuda())
        tr.step(l2, stats, labels)
        got = tr.get_centroids().cpu().numpy()
        lab = labels.cpu().numpy()
        touched = np.zeros(k, dtype=bool)
        mism = lab != ref["labels"]
        touched[lab[mism]] = True
        touched[ref["labels"][mism]] = True
        s = stats.cpu().numpy()
        assert int(s[1]) == ref["nsplit"]
        rel = np.linalg.norm(got - ref["centroids"], axis=1) / np.maximum(np.linalg.norm(ref["centroids"], axis=1), 1e-30)
        print(f"iter {it}: mismatched rows {int(mism.sum())}, touched {int(touched.sum())}, max rel (untouched) {rel[~touched].max():.2e}, obj {s[0]:.6g} vs {ref['obj']:.6g}")
        assert (rel[~touched] <= 1e-4).all()
        assert mism.mean() < 1e-4
        # the objective is evaluated from the exact cluster sums in double (sum |x|^2 + sum_j n_j |c_j|^2 - 2 <s_j, c_j>)
        assert abs(s[0] - ref["obj"]) <= 2e-5 * abs(ref["obj"])
        cents = ref["centroids"]


def test_split_clusters_on_device_matches_oracle():
    """Force empty clusters: duplicate centroids lose every tie to the lower index."""
    import torch
    from at_b200 import LloydTrainer
    from oracle import faiss_ref

    _, l2 = _frames(10)
    x = l2.cpu().numpy()
    k = 32
    cents = x[faiss_ref.rand_perm(x.shape[0], 1235)[:k]].copy()
    cents[5] = cents[2]
    cents[17] = cents[2]
    cents[31] = cents[30]
    ref = faiss_ref.lloyd_step(x, cents, exact=True)
    assert ref["nsplit"] == 3
    tr = LloydTrainer(64, k)
    tr.begin(l2)
    tr.set_centroids(torch.from_numpy(cents).cuda())
    stats = torch.zeros(4, device="cuda")
    tr.step(l2, stats)
    s = stats.cpu().numpy()
    assert int(s[1]) == 3 and int(s[3]) == 3
    got = tr.get_centroids().cpu().numpy()
    np.testing.assert_allclose(got, ref["centroids"], rtol=1e-4, atol=1e-7)
    assert abs(s[2] - faiss_ref.imbalance_factor(ref["counts"])) < 1e-3


def test_kmeans_faiss_api_free_running():
    """Tier (ii): same seeded subsample / init / niter as the oracle; report centroid agreement."""
    from at_b200 import Kmeans
    from oracle import faiss_ref

    _, l2 = _frames(40)
    x = l2.cpu().numpy()  # 40 clips * 87 frames = 3480 rows
    k = 8                 # 3480 > 8 * 256 -> FAISS subsamples to 2048 rows
    km = Kmeans(64, k, niter=10, verbose=False, gpu=True)
    obj = km.train(x)
    ref = faiss_ref.Kmeans(64, k, niter=10)
    ref.exact_search = True
    robj = ref.train(x)
    assert km.centroids.shape == (k, 64) and km.centroids.dtype == np.float32
    assert len(km.iteration_stats) == 10 and obj == km.obj[-1]
    rel = np.linalg.norm(km.centroids - ref.centroids, axis=1) / np.linalg.norm(ref.centroids, axis=1)
    print("free-running: fraction of centroids within 1e-4:", float((rel <= 1e-4).mean()), "rel obj diff", abs(obj - robj) / robj)
    assert (rel <= 1e-4).mean() >= 0.75
    assert abs(obj - robj) <= 1e-3 * robj
    # continuing from given centroids (the reference's second 10k-file batch, cluster_creator.py:55-56)
    km2 = Kmeans(64, k, niter=1)
    km2.train(x, init_centroids=km.centroids)
    ref2 = faiss_ref.Kmeans(64, k, niter=1)
    ref2.exact_search = True
    ref2.train(x, init_centroids=km.centroids)
    np.testing.assert_allclose(km2.centroids, ref2.centroids, rtol=1e-4, atol=1e-7)
    with pytest.raises(AttributeError):
        Kmeans(64, k, bogus=1)
    with pytest.raises(RuntimeError):
        Kmeans(64, 100).train(x[:10])
    bad = x.copy()
    bad[0, 0] = np.inf
    with pytest.raises(RuntimeError):
        Kmeans(64, k).train(bad)


def test_bincount():
    import torch
    from at_b200 import _lib

    lab = torch.randint(0, 300, (100000,), dtype=torch.int32, device="cuda")
    counts = torch.empty(300, dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().at_bincount(_lib.ptr(lab), lab.numel(), 300, _lib.ptr(counts), _lib.stream_ptr()))
    assert torch.equal(counts, torch.bincount(lab.long(), minlength=300))


# ------------------------------------------------------------------------------------------------ tcgen05 path
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("k,n", [(256, 20000), (1024, 30011), (500, 4097), (16, 300), (2000, 12800)])
def test_search_tensor_matches_exact_simt(k, n, mode):
    """The tcgen05 kernel re-checks near-tied top-2 candidates with the canonical fp32 formula, so labels must equal
    the exact SIMT kernel's (ties below fp32 resolution are the documented exception); distances agree to ~2e-5."""
    import torch
    from at_b200 import FlatL2, _lib

    spec, l2 = _frames(80, 220500)
    assert l2.shape[0] >= n
    x = l2[:n].contiguous()
    g = torch.Generator().manual_seed(k)
    c = l2[torch.randperm(l2.shape[0], generator=g)[:k].cuda()].contiguous()
    ix = FlatL2(64)
    ix.set_tc_mode(mode)  # 1: streamed operand tiles, 2: resident / K-sliced
    ix.set_centroids(c)
    ls, ds = ix.search(x, algo=_lib.ALGO_SIMT)
    lt, dt = ix.search(x, algo=_lib.ALGO_TENSOR)
    mism = (ls != lt)
    print(f"tensor vs simt k={k} n={n} mode={mode}: label mismatches {int(mism.sum())}")
    assert int(mism.sum()) <= max(1, n // 100000)
    # rows whose runner-up is safely behind skip the fp32 re-check: distance read off the accumulator
    torch.testing.assert_close(dt[~mism], ds[~mism], rtol=1e-4, atol=5e-6)
    # fused row normalisation
    lt2, dt2 = ix.search(spec[:n].contiguous(), l2norm_rows=True, algo=_lib.ALGO_TENSOR)
    ls2, ds2 = ix.search(spec[:n].contiguous(), l2norm_rows=True, algo=_lib.ALGO_SIMT)
    assert int((lt2 != ls2).sum()) <= max(1, n // 100000)
    assert torch.equal(ls2, ls)
    # oracle parity gate
    _check_labels(lt.cpu().numpy(), x.cpu().numpy(), c.cpu().numpy(), f"tensor k={k}")


def test_search_tensor_random_and_out_of_range_rows():
    import torch
    from at_b200 import FlatL2, _lib

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(5000, 64, device="cuda", generator=g)
    c = torch.randn(300, 64, device="cuda", generator=g)
    x[17] *= 1e4    # beyond the fp16 operand range -> exact fallback inside the kernel
    x[4000] *= 1e-6
    x[123] = 0
    ix = FlatL2(64)
    ix.set_centroids(c)
    ls, ds = ix.search(x, algo=_lib.ALGO_SIMT)
    lt, dt = ix.search(x, algo=_lib.ALGO_TENSOR)
    assert torch.equal(ls, lt)
    torch.testing.assert_close(dt, ds, rtol=1e-4, atol=1e-4)
    _check_labels(lt.cpu().numpy(), x.cpu().numpy(), c.cpu().numpy(), "tensor randn")


def test_search_tensor_adversarial_near_ties():
    """Near ties on purpose: every centroid has a twin that is bit-identical (exact tie: the lower index must win) or
    differs by one or a few ulps in a few coordinates (the fp32 formula decides), twins placed in the same and in different
    128-centroid tiles; rows sit on and next to the centroids.  The tensor path must return the exact kernel's labels."""
    import torch
    from at_b200 import FlatL2, _lib, row_l2norm

    g = torch.Generator(device="cuda").manual_seed(77)
    base = row_l2norm(torch.rand(300, 64, device="cuda", generator=g) + 0.2)
    twins = base.clone()
    # a third exact copies, a third one-ulp nudges of 3 coordinates, a third 1e-6 relative noise
    ulp = torch.nextafter(twins[100:200, :3], torch.full_like(twins[100:200, :3], 2.0))
    twins[100:200, :3] = ulp
    twins[200:] = twins[200:] * (1.0 + 1e-6 * torch.randn(100, 64, device="cuda", generator=g))
    for order in ("adjacent", "far"):
        if order == "adjacent":   # twin right after its original: same tile, same 16-column group most of the time
            c = torch.stack([base, twins], dim=1).reshape(-1, 64).contiguous()
        else:                     # twins 300 columns later: other tiles
            c = torch.cat([base, twins]).contiguous()
        x = torch.cat([base, twins, row_l2norm(base + 1e-4 * torch.randn(300, 64, device="cuda", generator=g)),
                       row_l2norm(0.5 * (base + twins))]).contiguous()
        ix = FlatL2(64)
        ix.set_centroids(c)
        ls, ds = ix.search(x, algo=_lib.ALGO_SIMT)
        lt, dt = ix.search(x, algo=_lib.ALGO_TENSOR)
        re, full = ix.tc_stats()
        print(f"{order}: tensor vs exact mismatches {int((ls != lt).sum())} of {x.shape[0]}, re-checked {re}, exact scans {full}")
        assert torch.equal(ls, lt), order
        assert torch.equal(ds, dt), order    # uncertified rows carry the canonical distance
        # exact copies: the lower index of the pair must have won for the rows sitting on them
        first = ls[:100]
        want = (torch.arange(100, device="cuda") * 2) if order == "adjacent" else torch.arange(100, device="cuda")
        assert torch.equal(first.long(), want)


def test_search_tensor_on_badly_centred_data():
    """The tensor path works on rows and centroids shifted by the centroid mean (the fp16 rounding errors then scale with
    |x - m| |c - m|).  The shift is only a conditioning device: data for which it is useless or harmful -- two far-apart
    blobs (the mean lies between them, far from every row), rows offset by a large constant, a single centroid cluster far
    from the rows -- must still give the exact kernel's labels (certified, re-checked or scanned exactly)."""
    import torch
    from at_b200 import FlatL2, _lib

    g = torch.Generator(device="cuda").manual_seed(21)
    n, k = 20000, 512
    blob = (torch.arange(n, device="cuda") % 2).float().reshape(-1, 1) * 40.0 - 20.0       # rows at -20 / +20
    x1 = blob + torch.randn(n, 64, device="cuda", generator=g)
    c1 = x1[torch.randperm(n, device="cuda", generator=g)[:k]].contiguous() + 0.05 * torch.randn(k, 64, device="cuda", generator=g)
    x2 = 300.0 + torch.rand(n, 64, device="cuda", generator=g)                              # large common offset
    c2 = x2[torch.randperm(n, device="cuda", generator=g)[:k]].contiguous() + 0.01
    x3 = torch.rand(n, 64, device="cuda", generator=g)                                      # centroids far from the rows
    c3 = 5.0 + 0.01 * torch.randn(k, 64, device="cuda", generator=g)
    for name, x, c in (("blobs", x1, c1), ("offset", x2, c2), ("far", x3, c3)):
        ix = FlatL2(64)
        ix.set_centroids(c)
        ls, _ = ix.search(x, algo=_lib.ALGO_SIMT)
        lt, _ = ix.search(x, algo=_lib.ALGO_TENSOR)
        mism = int((ls != lt).sum())
        re, full = ix.tc_stats()
        print(f"{name}: mismatches {mism}, re-checked {re}, exact scans {full}")
        # (no oracle gate here: with |x|^2 >> |x - c|^2 the expanded fp32 formula itself is ill-conditioned and two correct
        # evaluations of it differ in their summation order; the claim under test is tensor path == exact kernel)
        assert mism == 0, name


def test_lloyd_tensor_image_centre_survives_moving_centroids():
    """k-means builds the row image once, centred on the mean of the INITIAL centroids; the centroids then move away from it
    (here: initial centroids taken from one corner of the data).  Labels stay the exact kernel's in every iteration."""
    import torch
    from at_b200 import LloydTrainer, _lib

    g = torch.Generator(device="cuda").manual_seed(3)
    n, k = 30000, 96
    x = torch.cat([torch.randn(n // 2, 64, device="cuda", generator=g) * 0.3,
                   4.0 + torch.randn(n // 2, 64, device="cuda", generator=g) * 0.3]).contiguous()
    init = x[:k].contiguous()   # every initial centroid in the first blob
    outs = []
    for algo in (_lib.ALGO_SIMT, _lib.ALGO_TENSOR):
        tr = LloydTrainer(64, k, algo=algo)
        tr.begin(x)
        tr.set_centroids(init)
        labels = torch.empty(n, dtype=torch.int32, device="cuda")
        traj = []
        for _ in range(6):
            tr.step(x, None, labels)
            traj.append((labels.clone(), tr.get_centroids()))
        outs.append(traj)
    for (la, ca), (lb, cb) in zip(*outs):
        assert torch.equal(la, lb) and torch.equal(ca, cb)


def test_lloyd_tensor_vs_simt_bit_identical_centroids():
    """Same labels + exact integer sums => the two search paths give bit-identical k-means trajectories."""
    import torch
    from at_b200 import LloydTrainer, _lib

    _, l2 = _frames(60)
    k = 128
    init = l2[:: l2.shape[0] // k][:k].contiguous()
    outs = []
    for algo in (_lib.ALGO_SIMT, _lib.ALGO_TENSOR):
        tr = LloydTrainer(64, k, algo=algo)
        tr.begin(l2)
        tr.set_centroids(init)
        st = torch.zeros(5, 4, device="cuda")
        for it in range(5):
            tr.step(l2, st[it])
        outs.append((tr.get_centroids(), st.clone()))
    assert torch.equal(outs[0][0], outs[1][0])           # centroids: bit-identical
    assert torch.equal(outs[0][1][:, 1:], outs[1][1][:, 1:])  # nsplit, imbalance, empties
    torch.testing.assert_close(outs[0][1][:, 0], outs[1][1][:, 0], rtol=1e-4, atol=0)  # objective


@pytest.mark.parametrize("k", [16, 128, 1000])
def test_lloyd_incremental_update_bit_identical_to_full_regroup(k):
    """The incremental update (rows whose label changed leave one exact integer sum and join another) and the
    full regroup of every row give bit-identical centroids, counts-derived stats and labels at every iteration."""
    import torch
    from at_b200 import LloydTrainer

    _, l2 = _frames(120)
    n = l2.shape[0]
    init = l2[:: n // k][:k].contiguous()
    outs = []
    for inc in (False, True):
        tr = LloydTrainer(64, k)
        tr.set_incremental(inc)
        tr.begin(l2)
        tr.set_centroids(init)
        st = torch.zeros(8, 4, device="cuda")
        labs = torch.empty(8, n, dtype=torch.int32, device="cuda")
        cents = []
        for it in range(8):
            tr.step(l2, st[it], labs[it])
            cents.append(tr.get_centroids())
        outs.append((torch.stack(cents), st.clone(), labs))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][1], outs[1][1])
    changed = (outs[1][2][1:] != outs[1][2][:-1]).sum(dim=1)
    assert int(changed[0]) > 0   # the incremental path had rows to move


def test_kmeans_with_exactly_k_points_copies_the_training_set():
    """faiss::Clustering::train_encoded's corner case nx == k: centroids = the training set in input order, whatever
    init_centroids says, one all-zero iteration stat."""
    from at_b200 import Kmeans
    from oracle import faiss_ref

    _, l2 = _frames(4)
    x = l2.cpu().numpy()[:48]
    km = Kmeans(64, 48, niter=5)
    obj = km.train(x, init_centroids=x[::-1][:7].copy())
    ref = faiss_ref.Kmeans(64, 48, niter=5)
    robj = ref.train(x, init_centroids=x[::-1][:7].copy())
    assert np.array_equal(km.centroids, x) and np.array_equal(ref.centroids, x)
    assert len(km.iteration_stats) == 1 == len(ref.iteration_stats) and obj == 0.0 == robj


def test_lloyd_trainer_notices_new_rows_at_the_same_address():
    """The row image and the incremental sums are cached per training set; a different batch in the same buffer (or an
    in-place update) must not be searched through the stale image (at_kmeans_invalidate)."""
    import torch
    from at_b200 import LloydTrainer
    from oracle import faiss_ref

    _, l2 = _frames(30)
    a, b = l2[:1200].clone(), l2[1200:2400].clone()
    k = 64
    cents = a[torch.from_numpy(faiss_ref.rand_perm(1200, 1235)[:k].astype("int64")).cuda()].contiguous()

    def one_step(tr, x):
        tr.set_centroids(cents)
        tr.step(x, None)
        return tr.get_centroids()

    fresh = LloydTrainer(64, k)
    fresh.begin(b)
    want = one_step(fresh, b)
    tr = LloydTrainer(64, k)
    buf = a.clone()
    tr.begin(buf)
    one_step(tr, buf)
    tr.step(buf, None)                 # incremental state now describes batch a
    buf.copy_(b)                       # same pointer, same shape, new contents
    got = one_step(tr, buf)
    assert torch.equal(got, want)


def test_search_reusing_the_trainers_row_image_equals_the_plain_search():
    """Tokenizing the rows that were just clustered (at_index_search_trained_rows) = the ordinary search of the
    un-normalised rows with l2norm_rows, tensor and exact kernels alike."""
    import torch
    from at_b200 import FlatL2, LloydTrainer, _lib, row_l2norm

    spec, l2 = _frames(120)                      # 120 * 87 = 10,440 rows
    k = 256
    tr = LloydTrainer(64, k, algo=_lib.ALGO_TENSOR)
    tr.begin(l2)
    tr.set_centroids(l2[torch.randperm(l2.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(3))[:k]].contiguous())
    for _ in range(3):
        tr.step(l2, None)
    cents = row_l2norm(tr.get_centroids())
    ix = FlatL2(64)
    ix.set_centroids(cents)
    a, da = ix.search_trained_rows(tr, spec, l2norm_rows=True, want_dist=True)
    b, db = ix.search(spec, l2norm_rows=True, algo=_lib.ALGO_TENSOR)
    c, dc = ix.search(spec, l2norm_rows=True, algo=_lib.ALGO_SIMT)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(da, db) and torch.equal(da, dc)
    a2, _ = ix.search_trained_rows(tr, l2, l2norm_rows=False)          # the very array the trainer saw
    assert torch.equal(a2, ix.search(l2, algo=_lib.ALGO_SIMT)[0])
    d, _ = ix.search(spec, l2norm_rows=True, algo=_lib.ALGO_TENSOR)    # the index is back on its own scale afterwards
    assert torch.equal(d, c)
    with pytest.raises(RuntimeError):
        ix.search_trained_rows(tr, spec[:100].contiguous(), l2norm_rows=True)
