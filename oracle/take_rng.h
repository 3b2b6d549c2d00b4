// TEST INFRASTRUCTURE -- shared by the two oracle translation units only
// (oracle/take_oracle.cpp, oracle/ref_harness.cpp).  The product carries its own
// device copy in take_b200/csrc/rng.cuh; tests check the two agree.
//
// Counter-based sample streams.  Every path sample (pixel p, sample index s) owns
// an independent stream of 32-bit words W[0], W[1], ...:
//     W[j] = Philox4x32-10(counter = {j/4, p, s_lo, s_hi}, key = {seed_lo, seed_hi})[j%4]
// and the k-th `random_real` of the path consumes W[2k], W[2k+1] exactly the way
// libstdc++'s uniform_real_distribution<double>(mt19937) consumes two engine
// outputs (reference: src/take.h:89-91; <bits/random.tcc> generate_canonical):
//     r = (double(W[2k]) + double(W[2k+1]) * 2^32) / 2^64 ;  r >= 1 -> nextafter(1,0)
// so the unmodified reference integrators can be driven with the same numbers by
// pre-loading an mt19937 whose tempered outputs are W[] (see ref_harness.cpp).
#pragma once
#include <stdint.h>

static inline void take_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 4 consecutive stream words W[4*block .. 4*block+3] of sample (pixel, sample).
static inline void take_stream_block(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {block, pixel, (uint32_t)sample, (uint32_t)(sample >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    take_philox4x32_10(ctr, key, out);
}

static inline double take_words_to_real(uint32_t w0, uint32_t w1) {
    double sum = (double)w0 + (double)w1 * 4294967296.0;
    double r = sum / 18446744073709551616.0;
    if (r >= 1.0) r = 0x1.fffffffffffffp-1;  // nextafter(1.0, 0.0)
    return r;
}

// k-th random_real of the stream.
static inline double take_stream_real(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t k) {
    uint32_t w[4];
    take_stream_block(seed, pixel, sample, k >> 1, w);
    return take_words_to_real(w[2 * (k & 1)], w[2 * (k & 1) + 1]);
}
