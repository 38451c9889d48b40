// rfi_host.cpp -- (plain C++, built with g++ -O3; no CUDA)
// rfi_host -- host-side helper of the create_dataset path: the shuffle.
//
// preprocessor.py:758-763 draws ONE np.random.permutation(n_kept) from NumPy's global legacy
// generator between the two GPU phases, i.e. on the critical path while the GPU idles.
// NumPy's legacy shuffle costs ~30 ns per element (1.5 ms for 45 k patches); this is the same
// algorithm -- MT19937, masked-rejection `random_interval` with 32-bit draws, Fisher-Yates from
// the top (numpy/random/mtrand.pyx `_shuffle_raw`, numpy/random/src/distributions/
// distributions.c `random_interval`, numpy/random/src/mt19937/mt19937.c) -- run on a copy of
// the generator state, which the Python layer reads with np.random.get_state() and writes back
// with np.random.set_state(), so the permutation AND the generator's stream position are
// identical to what the reference leaves behind.
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rfi_b200.h"

namespace {

constexpr int kN = 624, kM = 397;

// AVX2 clone picked at load time where the host CPU has it (the loops vectorise: the recurrence
// reads entries at distance +1, not yet rewritten, and +-M, far outside a vector)
__attribute__((target_clones("avx2", "default")))
void mt_regenerate(uint32_t* mt) {
    constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
    int kk = 0;
    // the recurrence only reads entries at distance +1 (not yet rewritten) and +-(M) -- far
    // outside a vector -- so the loops may be vectorised
    for (; kk < kN - kM; ++kk) {
        const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
        mt[kk] = mt[kk + kM] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
    }
    for (; kk < kN - 1; ++kk) {
        const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
        mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
    }
    const uint32_t y = (mt[kN - 1] & kUpper) | (mt[0] & kLower);
    mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
}

__attribute__((target_clones("avx2", "default")))
void mt_temper(const uint32_t* __restrict__ key, uint32_t* __restrict__ out, int from) {
    for (int k = from; k < kN; ++k) {
        uint32_t y = key[k];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        out[k] = y;
    }
}

struct Mt {
    uint32_t* key;
    int pos;
    inline uint32_t next32() {
        if (pos == kN) { mt_regenerate(key); pos = 0; }
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline uint64_t next64() {  // mt19937_next64: high word first
        const uint64_t hi = next32();
        return (hi << 32) | next32();
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// The swap partner of every position, in draw order: part[t] is the partner of position
// i = n - 1 - t (t = 0 .. n - 2), i.e. the first masked draw v <= i of the stream at that point
// (numpy `random_interval`: 32-bit draws, mask = smallest 2^b - 1 >= i, rejection).
//
// While i stays inside (mask >> 1, mask] the mask is constant and the only loop-carried state is
// i itself.  The AVX2 path settles eight draws at once: lane l sees i_l in [i - l, i], so
// v_l <= i - l is accepted whatever happened before it and v_l > i is rejected whatever happened;
// only i - l < v_l <= i is ambiguous (probability ~ l / 2^b) and sends that block of eight through
// the scalar loop.  Accepted values are compacted with a 256-entry permutation table.
static inline void draws_scalar(const uint32_t* tb, int& k, int kend, uint32_t mask, uint32_t low,
                                uint32_t& i, uint32_t* part, uint32_t& t) {
    for (; k < kend && i > low; ++k) {
        const uint32_t v = tb[k] & mask;
        part[t] = v;
        const uint32_t ok = (v <= i) ? 1u : 0u;
        t += ok;
        i -= ok;
    }
}

#if defined(__x86_64__)
#include <immintrin.h>
alignas(32) static uint32_t g_compact_lut[256][8];
static bool g_lut_ready = false;
static void build_lut() {
    for (int m = 0; m < 256; ++m) {
        int n = 0;
        for (int b = 0; b < 8; ++b) if (m & (1 << b)) g_compact_lut[m][n++] = (uint32_t)b;
        for (; n < 8; ++n) g_compact_lut[m][n] = 0;
    }
    g_lut_ready = true;
}
__attribute__((target("avx2,popcnt")))
static void draws_avx2(const uint32_t* tb, int& k, int kend, uint32_t mask, uint32_t low,
                       uint32_t& i, uint32_t* part, uint32_t& t) {
    const __m256i vmask = _mm256_set1_epi32((int)mask);
    const __m256i lanes = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    // all quantities are < 2^31 (n <= 2^31 - 1), so signed compares order them correctly.
    // Super-blocks of 32 draws: the four vectors are judged against the SAME i (draw p sees
    // i_p in [i - p, i]), so their compares do not wait for each other's accept counts -- the
    // loop-carried chain through i is paid once per 32 draws.  Worth it while ambiguity
    // (~ 32 * 16 / 2^b per super-block) is rare; below that, blocks of 8.
    while (mask >= 0x3fffu && k + 32 <= kend && i > low + 32) {
        const __m256i vi = _mm256_set1_epi32((int)i);
        __m256i v[4];
        int acc[4], amb = 0;
#pragma GCC unroll 4
        for (int b = 0; b < 4; ++b) {
            v[b] = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(tb + k + 8 * b)), vmask);
            const __m256i rej = _mm256_cmpgt_epi32(v[b], vi);                                               // v > i
            const __m256i off = _mm256_add_epi32(lanes, _mm256_set1_epi32(8 * b));
            const __m256i notsure = _mm256_cmpgt_epi32(v[b], _mm256_sub_epi32(vi, off));                     // v > i - p
            amb |= _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_andnot_si256(rej, notsure)));
            acc[b] = (~_mm256_movemask_ps(_mm256_castsi256_ps(notsure))) & 0xff;
        }
        if (amb) {  // rare: settle these 32 one by one
            int kk = k;
            draws_scalar(tb, kk, k + 32, mask, low, i, part, t);
            k = kk;
            continue;
        }
        uint32_t tt = t;
#pragma GCC unroll 4
        for (int b = 0; b < 4; ++b) {
            const __m256i idx = _mm256_load_si256(reinterpret_cast<const __m256i*>(g_compact_lut[acc[b]]));
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(part + tt), _mm256_permutevar8x32_epi32(v[b], idx));
            tt += (uint32_t)__builtin_popcount((unsigned)acc[b]);
        }
        i -= tt - t;
        t = tt;
        k += 32;
    }
    while (k + 8 <= kend && i > low + 8) {
        const __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(tb + k)), vmask);
        const __m256i vi = _mm256_set1_epi32((int)i);
        const __m256i rej = _mm256_cmpgt_epi32(v, vi);                              // v > i
        const __m256i notsure = _mm256_cmpgt_epi32(v, _mm256_sub_epi32(vi, lanes));  // v > i - l
        const int amb = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_andnot_si256(rej, notsure)));
        if (amb) {  // rare: settle these eight one by one
            int kk = k;
            draws_scalar(tb, kk, k + 8, mask, low, i, part, t);
            k = kk;
            continue;
        }
        const int acc = (~_mm256_movemask_ps(_mm256_castsi256_ps(notsure))) & 0xff;
        const __m256i idx = _mm256_load_si256(reinterpret_cast<const __m256i*>(g_compact_lut[acc]));
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(part + t), _mm256_permutevar8x32_epi32(v, idx));
        const uint32_t c = (uint32_t)__builtin_popcount((unsigned)acc);
        t += c;
        i -= c;
        k += 8;
    }
}
#endif

// grow-only per-thread scratch: a fresh malloc of a few hundred KB per call would be an mmap plus
// its page faults on the critical path between the two GPU phases
template <typename V>
static V* scratch(int which, size_t count) {
    static thread_local void* buf[1] = {nullptr};
    static thread_local size_t cap[1] = {0};
    const size_t bytes = count * sizeof(V);
    if (cap[which] < bytes) {
        free(buf[which]);
        cap[which] = bytes + bytes / 2 + 64;
        buf[which] = malloc(cap[which]);
        if (!buf[which]) cap[which] = 0;
    }
    return static_cast<V*>(buf[which]);
}

// Fills part[0 .. n - 2] (room for 8 more entries required) and advances the generator.
static void legacy_partners(uint32_t* mt_key, int32_t* mt_pos, uint32_t n, uint32_t* part) {
    uint32_t tb[kN];
    int pos = *mt_pos;
    mt_temper(mt_key, tb, pos);
#if defined(__x86_64__)
    const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt");
    if (avx2 && !g_lut_ready) build_lut();
#endif
    uint32_t i = n - 1, t = 0;
    while (i >= 1) {
        if (pos == kN) { mt_regenerate(mt_key); mt_temper(mt_key, tb, 0); pos = 0; }
        int k = pos;
        while (k < kN && i >= 1) {
            const uint32_t mask = 0xffffffffu >> __builtin_clz(i);  // smallest 2^b - 1 >= i
            const uint32_t low = mask >> 1;                         // i stays in (low, mask]
#if defined(__x86_64__)
            if (avx2) draws_avx2(tb, k, kN, mask, low, i, part, t);
#endif
            // the tail of the segment / of the generator block, or everything without AVX2
            draws_scalar(tb, k, kN, mask, low, i, part, t);
        }
        pos = k;
    }
    *mt_pos = pos;
}

// RandomState.shuffle of a 1-D int64 array (numpy/random/mtrand.pyx `_shuffle_raw`)
static int legacy_shuffle(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int64_t* out) {
    if (!mt_key || !mt_pos || n < 0 || (n > 0 && !out) || *mt_pos < 0 || *mt_pos > kN) return RFI_E_INVALID;
    if (n > 0x7fffffffLL) {  // 64-bit draws above 2^32 - 1: the plain loop
        Mt g{mt_key, *mt_pos};
        for (int64_t i = n - 1; i >= 1; --i) {
            const uint64_t max = (uint64_t)i;
            uint64_t mask = max, value;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
            mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
            if (max <= 0xffffffffull) {
                while ((value = (g.next32() & mask)) > max) {}
            } else {
                while ((value = (g.next64() & mask)) > max) {}
            }
            const int64_t t = out[value];
            out[value] = out[i];
            out[i] = t;
        }
        *mt_pos = g.pos;
        return RFI_OK;
    }
    if (n < 2) return RFI_OK;
    // phase 1: the swap partner of every position; phase 2: the swaps.  Splitting them keeps the
    // RNG chain and the memory chain apart.
    uint32_t* part = scratch<uint32_t>(0, (size_t)n + 8);
    if (!part) return RFI_E_INVALID;
    legacy_partners(mt_key, mt_pos, (uint32_t)n, part);
    for (uint32_t q = (uint32_t)(n - 1), t = 0; q >= 1; --q, ++t) {
        const uint32_t j = part[t];
        const int64_t a = out[q], b = out[j];
        out[j] = a;
        out[q] = b;
    }
    return RFI_OK;
}

extern "C" int rfi_legacy_permutation(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int64_t* out) {
    if (n < 0 || (n > 0 && !out)) return RFI_E_INVALID;
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    return legacy_shuffle(mt_key, mt_pos, n, out);
}

// Blank-patch removal, shuffle and truncation of preprocessor.py:746-763 + :356-359 in one
// host call: per-group flag counts -> canonical order of the kept patches -> legacy shuffle ->
// destination slot of every patch.
extern "C" int rfi_plan_slots(const rfi_plan_t* plan, const int32_t* n_flagged, int64_t stride_bytes,
                              int shuffle, uint32_t* mt_key, int32_t* mt_pos, int64_t num_patches,
                              int64_t* order, int64_t* dest, int64_t* n_out) {
    if (!plan || !order || !dest || !n_out || plan->patch <= 0) return RFI_E_INVALID;
    const int64_t C = plan->channels, T = plan->times, P = plan->patch;
    const int R = plan->rotations;
    const bool skip = C <= P && T <= P;
    const int64_t nhc = skip ? 1 : (C + P - 1) / P, nwc = skip ? 1 : (T + P - 1) / P;
    const bool padded = !skip && (C % P || T % P);
    const int64_t per = nhc * nwc, W = plan->n_waterfalls;
    const int64_t n0 = W * R * per;
    auto flagged = [&](int64_t g) {
        return *reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(n_flagged) + g * stride_bytes) > 0;
    };
    int64_t cnt = 0;
    if (!shuffle || !n_flagged) {  // inference mode: canonical order, nothing dropped (:345-353)
        cnt = (num_patches > 0 && num_patches < n0) ? num_patches : n0;  // :356-359
        for (int64_t q = 0; q < n0; ++q) dest[q] = q < cnt ? q : -1;
        for (int64_t q = 0; q < cnt; ++q) order[q] = q;
        *n_out = cnt;
        return RFI_OK;
    } else {
        int64_t q = 0;
        for (int64_t w = 0; w < W; ++w) {
            for (int r = 0; r < R; ++r) {
                if (padded) {  // one statistic group per patch, canonical order
                    for (int64_t t = 0; t < per; ++t, ++q)
                        if (flagged(q)) order[cnt++] = q;
                    continue;
                }
                const int64_t gb = w * per;
                if (r <= 1) {  // rotated grid nhc x nwc; r = 1 lists the row blocks reversed
                    for (int64_t bi = 0; bi < nhc; ++bi) {
                        const int64_t ti = r == 0 ? bi : nhc - 1 - bi;
                        for (int64_t bj = 0; bj < nwc; ++bj, ++q)
                            if (flagged(gb + ti * nwc + bj)) order[cnt++] = q;
                    }
                } else {  // transposed grid nwc x nhc; r = 3 lists its row blocks reversed
                    for (int64_t bi = 0; bi < nwc; ++bi) {
                        const int64_t tj = r == 2 ? bi : nwc - 1 - bi;
                        for (int64_t bj = 0; bj < nhc; ++bj, ++q)
                            if (flagged(gb + bj * nwc + tj)) order[cnt++] = q;
                    }
                }
            }
        }
        if (cnt == 0) {  // no flagged patch anywhere: keep all (:752-756)
            for (int64_t i = 0; i < n0; ++i) order[i] = i;
            cnt = n0;
        }
        const int rc = legacy_shuffle(mt_key, mt_pos, cnt, order);  // kept[permutation(n)] == shuffle(kept)
        if (rc) return rc;
        if (num_patches > 0 && num_patches < cnt) cnt = num_patches;  // :356-359
        for (int64_t i = 0; i < n0; ++i) dest[i] = -1;
        for (int64_t k = 0; k < cnt; ++k) dest[order[k]] = k;
    }
    *n_out = cnt;
    return RFI_OK;
}
