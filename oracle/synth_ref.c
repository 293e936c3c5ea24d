/*
 * oracle/synth_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * C twin of oracle/synth_ref.py (and therefore of the device generator k_synth): the same integer-only arithmetic,
 * OpenMP over clips, so the CPU arm of bench.py can build its 16-bit-PCM-valued input clips in seconds without
 * touching the product library.  tests/test_oracle_synth.py checks it bit for bit against the numpy version.
 */
#include <stdint.h>

static uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

/* out: count x n_samples floats (int16 / 32768); pcm (may be NULL): the same samples as int16. */
void ref_synth_clips(uint32_t seed, int64_t first, int64_t count, int64_t n_samples, const int16_t *table,
                     float *out, int16_t *pcm) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t clip = 0; clip < count; clip++) {
        uint32_t h0 = hash32(seed * 0x9E3779B1u + (uint32_t)(first + clip));
        h0 = hash32(h0 ^ 0x85EBCA6Bu);
        const int P = 3 + (int)(hash32(h0 + 1u) % 6u);
        const int na = 16 + (int)(hash32(h0 + 2u) % 240u);
        const int gain = 96 + (int)(hash32(h0 + 3u) % 160u);
        uint32_t hp[8], inc[8], phi0[8];
        int amp[8];
        for (int p = 0; p < P; p++) {
            hp[p] = hash32(h0 + 16u + (uint32_t)p);
            const uint32_t hq = hash32(hp[p]);
            inc[p] = (15600000u + ((hq >> 8) & 0xFFFFFFu)) << (hp[p] % 7u);
            phi0[p] = hash32(hp[p] + 0x1234567u);
            amp[p] = 512 + (int)(hash32(hq + 7u) % 3584u);
        }
        for (int64_t n = 0; n < n_samples; n++) {
            const uint32_t un = (uint32_t)n, seg = un >> 13;
            const int r = (int)(un & 8191u);
            int total = 0;
            for (int p = 0; p < P; p++) {
                const uint32_t phase = phi0[p] + un * inc[p];
                const int s = table[phase >> 20];
                int e0 = (int)(hash32((hp[p] ^ (seg * 0x9E3779B1u)) + 0x55u) % 384u) - 128;
                int e1 = (int)(hash32((hp[p] ^ ((seg + 1u) * 0x9E3779B1u)) + 0x55u) % 384u) - 128;
                e0 = e0 < 0 ? 0 : e0;
                e1 = e1 < 0 ? 0 : e1;
                const int env = (e0 * (8192 - r) + e1 * r) >> 13;
                total += (((s * amp[p]) >> 12) * env) >> 8;
            }
            total = (total * gain) >> 9;
            const uint32_t hn = hash32(h0 ^ hash32(un + 0x68E31DA4u));
            const int nz = (int)(hn % (uint32_t)(2 * na + 1)) - na;
            int v = total + nz;
            v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
            if (out) out[clip * n_samples + n] = (float)v * (1.0f / 32768.0f);
            if (pcm) pcm[clip * n_samples + n] = (int16_t)v;
        }
    }
}
