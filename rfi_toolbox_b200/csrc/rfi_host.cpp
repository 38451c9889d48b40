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
    // Same draws, same swaps, restructured for the host core: the 624 outputs of a generator
    // block are tempered in one vectorisable loop, and a rejected draw (value > i) becomes a
    // swap of out[i] with itself, so the only branches left are the loop exits.
    uint32_t tb[kN];
    int pos = *mt_pos;
    auto temper_from = [&](int from) { mt_temper(mt_key, tb, from); };
    temper_from(pos);
    // phase 1: the swap partner of every position (rejected draws are simply overwritten);
    // phase 2: the swaps.  Splitting them keeps the RNG chain and the memory chain apart.
    if (n < 2) return RFI_OK;
    uint32_t* js = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (size_t)n));
    if (!js) return RFI_E_INVALID;
    uint32_t i = (uint32_t)(n - 1);
    while (i >= 1) {
        if (pos == kN) { mt_regenerate(mt_key); temper_from(0); pos = 0; }
        int k = pos;
        // i stays inside (mask >> 1, mask] for a whole inner loop, so the mask is a loop constant
        // and the only loop-carried chain is compare -> subtract
        while (k < kN && i >= 1) {
            const uint32_t mask = 0xffffffffu >> __builtin_clz(i);  // smallest 2^b - 1 >= i
            const uint32_t low = mask >> 1;
            for (; k < kN && i > low; ++k) {
                const uint32_t v = tb[k] & mask;
                js[i] = v;
                i -= (v <= i) ? 1u : 0u;
            }
        }
        pos = k;
    }
    for (uint32_t q = (uint32_t)(n - 1); q >= 1; --q) {
        const uint32_t j = js[q];
        const int64_t a = out[q], b = out[j];
        out[j] = a;
        out[q] = b;
    }
    free(js);
    *mt_pos = pos;
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
        for (int64_t q = 0; q < n0; ++q) order[q] = q;
        cnt = n0;
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
    }
    if (num_patches > 0 && num_patches < cnt) cnt = num_patches;  // :356-359
    for (int64_t i = 0; i < n0; ++i) dest[i] = -1;
    for (int64_t k = 0; k < cnt; ++k) dest[order[k]] = k;
    *n_out = cnt;
    return RFI_OK;
}
