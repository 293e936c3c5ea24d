"""Pins the FAISS restatement (oracle/faiss_ref.{c,py}) to known answers that need no FAISS binary
 and to fp64 brute force."""
import numpy as np
import pytest

from oracle import faiss_ref


def test_mt19937_known_answers():
    # computed with g++ 13.3 std::mt19937 and numpy MT19937
    assert faiss_ref.mt19937_words(1235, 5).tolist() == [4096379083, 3613105954, 4261150722, 3397365999, 2059878851]
    assert faiss_ref.rand_perm(10, 1234).tolist() == [5, 4, 8, 2, 6, 9, 1, 7, 0, 3]
    assert abs(float(faiss_ref.rand_floats(1234, 1)[0]) - 0.191519454) < 1e-8


@pytest.mark.parametrize("seed", [1234, 1235, 7])
def test_mt19937_matches_numpy_legacy(seed):
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    assert (faiss_ref.mt19937_words(seed, 2000) == bg.random_raw(2000).astype(np.uint32)).all()


def test_rand_perm_is_fisher_yates_of_the_stream():
    n, seed = 1000, 1234
    words = faiss_ref.mt19937_words(seed, n)
    perm = list(range(n))
    for i in range(n - 1):
        i2 = i + int(words[i]) % (n - i)
        perm[i], perm[i2] = perm[i2], perm[i]
    assert faiss_ref.rand_perm(n, seed).tolist() == perm


def _data(n, d, k, seed=0):
    rng = np.random.default_rng(seed)
    cent = rng.random((k, d), dtype=np.float32)
    x = cent[rng.integers(0, k, n)] + 0.05 * rng.standard_normal((n, d)).astype(np.float32)
    x = np.abs(x).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-10
    return x.astype(np.float32), cent


def test_search_matches_fp64_bruteforce_outside_near_ties():
    x, c = _data(5000, 64, 200)
    idx = faiss_ref.IndexFlatL2(64)
    idx.add(c)
    D, I = idx.search(x, 1)
    assert D.shape == (5000, 1) and I.shape == (5000, 1) and I.dtype == np.int64 and D.dtype == np.float32
    l64, d1, d2 = faiss_ref.assign_l2_f64(x, c)
    gap = (d2 - d1) / np.maximum(d1, 1e-30)
    bad = I[:, 0] != l64
    assert (gap[bad] < 1e-4).all()
    ls, s1, s2 = faiss_ref.assign_l2_scalar(x, c)
    bad = ls != l64
    assert (gap[bad] < 1e-4).all()
    np.testing.assert_allclose(D[:, 0], d1, rtol=2e-3, atol=2e-6)


def test_search_small_batch_uses_direct_distance():
    x, c = _data(7, 16, 9)
    idx = faiss_ref.IndexFlatL2(16)
    idx.add(c)
    D, I = idx.search(x, 1)
    ref = ((x[:, None, :].astype(np.float64) - c[None]) ** 2).sum(-1)
    assert (I[:, 0] == ref.argmin(1)).all()


def test_lowest_index_wins_exact_ties():
    c = np.zeros((4, 8), dtype=np.float32)
    c[1] = 1.0
    c[3] = 1.0  # duplicate of row 1
    x = np.ones((25, 8), dtype=np.float32)
    idx = faiss_ref.IndexFlatL2(8)
    idx.add(c)
    _, I = idx.search(x, 1)
    assert (I == 1).all()
    assert (faiss_ref.assign_l2_scalar(x, c)[0] == 1).all()


def test_compute_centroids_in_order_fp32_and_empty_cluster_zero():
    x, _ = _data(3000, 32, 10)
    assign = np.random.default_rng(1).integers(0, 9, 3000)  # cluster 9 empty
    cen, h = faiss_ref.compute_centroids(x, assign, 10)
    assert h[9] == 0 and (cen[9] == 0).all()
    for c in range(9):
        rows = x[assign == c]
        acc = np.zeros(32, dtype=np.float32)
        for r in rows:
            acc += r
        exp = acc * np.float32(1.0 / np.float32(len(rows)))
        assert np.array_equal(cen[c], exp)
        np.testing.assert_allclose(cen[c], rows.astype(np.float64).mean(0), rtol=1e-5)


def test_split_clusters_restatement():
    k, d, n = 8, 6, 1000
    cen = np.arange(k * d, dtype=np.float32).reshape(k, d) + 1
    h = np.array([300, 0, 200, 100, 0, 150, 250, 0], dtype=np.float32)
    out_c, out_h, nsplit = faiss_ref.split_clusters(cen, h, n)
    assert nsplit == 3
    # python re-derivation from the mt19937 stream
    r = faiss_ref.rand_floats(1234, 10000)
    ri = 0
    c2, h2 = cen.copy(), h.copy()
    for ci in range(k):
        if h2[ci] == 0:
            cj = 0
            while True:
                p = np.float32((np.float64(h2[cj]) - 1.0) / np.float64(np.float32(n - k)))
                rr = r[ri]
                ri += 1
                if rr < p:
                    break
                cj = (cj + 1) % k
            c2[ci] = c2[cj]
            for j in range(d):
                s = 1 + 1 / 1024.0 if j % 2 == 0 else 1 - 1 / 1024.0
                t = 1 - 1 / 1024.0 if j % 2 == 0 else 1 + 1 / 1024.0
                c2[ci, j] = np.float32(np.float64(c2[ci, j]) * s)
                c2[cj, j] = np.float32(np.float64(c2[cj, j]) * t)
            h2[ci] = h2[cj] / 2
            h2[cj] -= h2[ci]
    assert np.array_equal(out_c, c2) and np.array_equal(out_h, h2)
    assert out_h.sum() == h.sum()


def test_kmeans_api_and_semantics():
    x, _ = _data(6000, 16, 20, seed=3)
    km = faiss_ref.Kmeans(16, 20, niter=5, verbose=False, gpu=False)
    obj = km.train(x)
    assert km.centroids.shape == (20, 16) and km.centroids.dtype == np.float32
    assert len(km.iteration_stats) == 5 and obj == km.obj[-1]
    # 6000 > 20*256 -> subsampled to 5120 points with rand_perm(seed 1234); init = x_sub[rand_perm(5120, 1235)[:20]]
    perm = faiss_ref.rand_perm(6000, 1234)[:5120]
    xs = x[perm]
    init = xs[faiss_ref.rand_perm(5120, 1235)[:20]]
    c = init
    for _ in range(5):
        c = faiss_ref.lloyd_step(xs, c)["centroids"]
    np.testing.assert_array_equal(km.centroids, c)
    # objective non-increasing when nothing is split
    objs = km.obj
    if all(s["nsplit"] == 0 for s in km.iteration_stats):
        assert (np.diff(objs) <= 1e-3 * objs[0]).all()
    # continuing from init_centroids skips the random init
    km2 = faiss_ref.Kmeans(16, 20, niter=1)
    km2.train(x, init_centroids=km.centroids)
    exp = faiss_ref.lloyd_step(xs, km.centroids)["centroids"]
    np.testing.assert_array_equal(km2.centroids, exp)
    with pytest.raises(AttributeError):
        faiss_ref.Kmeans(16, 20, not_a_field=1)
    with pytest.raises(RuntimeError):
        faiss_ref.Kmeans(16, 20).train(x[:10])
    bad = x.copy()
    bad[3, 3] = np.nan
    with pytest.raises(RuntimeError):
        faiss_ref.Kmeans(16, 20).train(bad)
