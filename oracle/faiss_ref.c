/*
 * oracle/faiss_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the FAISS 1.8.0 routines that audio-tokens' k-means /
 * tokenization stages execute (reference call sites: processors/cluster_creator.py:42-58,
 * processors/spec_tokenizer.py:77,123-127).  FAISS itself (conda faiss-gpu=1.8.0,
 * environment.yml:69,137) is NOT vendored in /root/reference and cannot be installed in
 * this image, so this file restates the published algorithm of upstream
 * facebookresearch/faiss tag v1.8.0:
 *
 *   faiss/utils/random.cpp      RandomGenerator (std::mt19937), rand_int, rand_float, rand_perm
 *   faiss/Clustering.cpp        subsample_training_set, compute_centroids, split_clusters
 *   faiss/utils/distances.cpp   exhaustive_L2sqr_blas:  dis = |x|^2 + |y|^2 - 2<x,y>, clamp at 0,
 *                               strict '<' top-1 so the lowest index wins exact ties
 *
 * PARITY UNPINNED against real FAISS output (no FAISS binary here, the reference ships no
 * golden vectors).  Pinned instead by known answers computed with g++ 13.3 std::mt19937 and
 * numpy's MT19937: see tests/test_oracle_faiss.py.
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared, no -ffast-math: summation order matters)
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ mt19937 */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[i] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* faiss/utils/random.cpp: rand_float() = mt() / float(mt.max()) */
static float mt_rand_float(mt19937_t *g) {
    return (float)mt_next(g) / (float)0xFFFFFFFFu;
}

void ref_mt19937_words(int64_t seed, int n, uint32_t *out) {
    mt19937_t g;
    mt_seed(&g, (uint32_t)seed);
    for (int i = 0; i < n; i++) out[i] = mt_next(&g);
}

void ref_rand_floats(int64_t seed, int n, float *out) {
    mt19937_t g;
    mt_seed(&g, (uint32_t)seed);
    for (int i = 0; i < n; i++) out[i] = mt_rand_float(&g);
}

/* faiss/utils/random.cpp rand_perm: forward Fisher-Yates, i2 = i + mt() % (n - i) */
void ref_rand_perm(int32_t *perm, int64_t n, int64_t seed) {
    mt19937_t g;
    mt_seed(&g, (uint32_t)seed);
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    for (int64_t i = 0; i + 1 < n; i++) {
        int64_t i2 = i + (int64_t)((uint64_t)mt_next(&g) % (uint64_t)(n - i));
        int32_t t = perm[i];
        perm[i] = perm[i2];
        perm[i2] = t;
    }
}

/* ------------------------------------------------------------ assignment */
/* |x_i|^2 + |c_j|^2 - 2 <x_i, c_j>, fp32, sequential accumulation over d; strict '<'.
 * Also emits the runner-up distance so tests can apply the near-tie carve-out. */
void ref_assign_l2(const float *x, int64_t n, const float *c, int64_t k, int64_t d,
                   int64_t *labels, float *dist, float *dist2) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float *xi = x + i * d;
        float xn = 0.f;
        for (int64_t t = 0; t < d; t++) xn += xi[t] * xi[t];
        float best = INFINITY, second = INFINITY;
        int64_t bj = -1;
        for (int64_t j = 0; j < k; j++) {
            const float *cj = c + j * d;
            float cn = 0.f, ip = 0.f;
            for (int64_t t = 0; t < d; t++) {
                cn += cj[t] * cj[t];
                ip += xi[t] * cj[t];
            }
            float dis = xn + cn - 2.f * ip;
            if (dis < 0.f) dis = 0.f;
            if (dis < best) {
                second = best;
                best = dis;
                bj = j;
            } else if (dis < second) {
                second = dis;
            }
        }
        labels[i] = bj;
        if (dist) dist[i] = best;
        if (dist2) dist2[i] = second;
    }
}

/* fp64 exact argmin of sum (x-c)^2 with top-2 distances: the "truth" for gap analysis. */
void ref_assign_l2_f64(const float *x, int64_t n, const float *c, int64_t k, int64_t d,
                       int64_t *labels, double *dist, double *dist2) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float *xi = x + i * d;
        double best = INFINITY, second = INFINITY;
        int64_t bj = -1;
        for (int64_t j = 0; j < k; j++) {
            const float *cj = c + j * d;
            double s = 0.0;
            for (int64_t t = 0; t < d; t++) {
                double df = (double)xi[t] - (double)cj[t];
                s += df * df;
            }
            if (s < best) {
                second = best;
                best = s;
                bj = j;
            } else if (s < second) {
                second = s;
            }
        }
        labels[i] = bj;
        if (dist) dist[i] = best;
        if (dist2) dist2[i] = second;
    }
}

/* ------------------------------------------------------ compute_centroids */
/* faiss/Clustering.cpp compute_centroids (no weights, no codec, k_frozen = 0):
 * zero, add member rows in point-index order in fp32, hassign[c] += 1.0f, then
 * c *= 1 / hassign[c] for non-empty clusters (empty ones stay zero). */
void ref_compute_centroids(int64_t d, int64_t k, int64_t n, const float *x,
                           const int64_t *assign, float *hassign, float *centroids) {
    memset(centroids, 0, sizeof(float) * (size_t)(d * k));
    memset(hassign, 0, sizeof(float) * (size_t)k);
    /* like upstream: every thread walks all points and accumulates the clusters of its own slice
     * [k rank / nt, k (rank+1) / nt), so each cluster is summed by one thread in point-index order and the result
     * does not depend on the thread count */
#pragma omp parallel
    {
        int nt = 1, rank = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        rank = omp_get_thread_num();
#endif
        const int64_t c0 = (k * rank) / nt, c1 = (k * (rank + 1)) / nt;
        for (int64_t i = 0; i < n; i++) {
            int64_t ci = assign[i];
            if (ci >= c0 && ci < c1) {
                float *c = centroids + ci * d;
                const float *xi = x + i * d;
                hassign[ci] += 1.0f;
                for (int64_t j = 0; j < d; j++) c[j] += xi[j];
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < k; ci++) {
        if (hassign[ci] == 0) continue;
        float norm = 1 / hassign[ci];
        float *c = centroids + ci * d;
        for (int64_t j = 0; j < d; j++) c[j] *= norm;
    }
}

/* exhaustive_L2sqr_blas' epilogue for one (nx x ny) block of inner products ip (row stride ldip, produced by an sgemm):
 * dis = x_norm + y_norm - 2 ip, negative -> 0, strict '<' against the running best (the lowest index wins ties, blocks are
 * visited in ascending j0).  faiss/utils/distances.cpp. */
void ref_l2_block_argmin(const float *ip, int64_t nx, int64_t ny, int64_t ldip, const float *x_norms,
                         const float *y_norms, int64_t j0, float *best, int64_t *labels) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        const float *row = ip + i * ldip;
        const float xn = x_norms[i];
        float b = best[i];
        int64_t bj = labels[i];
        for (int64_t j = 0; j < ny; j++) {
            float dis = xn + y_norms[j] - 2 * row[j];
            if (dis < 0) dis = 0;
            if (dis < b) {
                b = dis;
                bj = j0 + j;
            }
        }
        best[i] = b;
        labels[i] = bj;
    }
}

/* --------------------------------------------------------- split_clusters */
/* faiss/Clustering.cpp split_clusters, EPS = 1/1024., fresh RandomGenerator(1234). */
int ref_split_clusters(int64_t d, int64_t k, int64_t n, float *hassign, float *centroids) {
    const double EPS = 1 / 1024.;
    int nsplit = 0;
    mt19937_t rng;
    mt_seed(&rng, 1234u);
    for (int64_t ci = 0; ci < k; ci++) {
        if (hassign[ci] == 0) {
            int64_t cj;
            for (cj = 0; 1; cj = (cj + 1) % k) {
                float p = (float)((hassign[cj] - 1.0) / (float)(n - k));
                float r = mt_rand_float(&rng);
                if (r < p) break;
            }
            memcpy(centroids + ci * d, centroids + cj * d, sizeof(float) * (size_t)d);
            for (int64_t j = 0; j < d; j++) {
                if (j % 2 == 0) {
                    centroids[ci * d + j] = (float)(centroids[ci * d + j] * (1 + EPS));
                    centroids[cj * d + j] = (float)(centroids[cj * d + j] * (1 - EPS));
                } else {
                    centroids[ci * d + j] = (float)(centroids[ci * d + j] * (1 - EPS));
                    centroids[cj * d + j] = (float)(centroids[cj * d + j] * (1 + EPS));
                }
            }
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
            nsplit++;
        }
    }
    return nsplit;
}
