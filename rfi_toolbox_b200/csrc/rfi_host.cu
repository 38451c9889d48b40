// rfi_host.cu -- host-side helper of the create_dataset path: the shuffle.
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

#include "../../include/rfi_b200.h"

namespace {

constexpr int kN = 624, kM = 397;

inline void mt_regenerate(uint32_t* mt) {
    constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
    int kk = 0;
    // the recurrence only reads entries at distance +1 (not yet rewritten) and +-(M) -- far
    // outside a vector -- so the loops may be vectorised
#pragma GCC ivdep
    for (; kk < kN - kM; ++kk) {
        const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
        mt[kk] = mt[kk + kM] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
    }
#pragma GCC ivdep
    for (; kk < kN - 1; ++kk) {
        const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
        mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
    }
    const uint32_t y = (mt[kN - 1] & kUpper) | (mt[0] & kLower);
    mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrixA);
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

extern "C" int rfi_legacy_permutation(uint32_t* mt_key, int32_t* mt_pos, int64_t n, int64_t* out) {
    if (!mt_key || !mt_pos || n < 0 || (n > 0 && !out) || *mt_pos < 0 || *mt_pos > kN) return RFI_E_INVALID;
    for (int64_t i = 0; i < n; ++i) out[i] = i;
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
    auto temper_from = [&](int from) {
        for (int k = from; k < kN; ++k) {
            uint32_t y = mt_key[k];
            y ^= (y >> 11);
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= (y >> 18);
            tb[k] = y;
        }
    };
    temper_from(pos);
    uint32_t i = n > 0 ? (uint32_t)(n - 1) : 0u;
    while (i >= 1) {
        if (pos == kN) { mt_regenerate(mt_key); temper_from(0); pos = 0; }
        int k = pos;
        uint32_t mask = 0xffffffffu >> __builtin_clz(i);  // smallest 2^b - 1 >= i
        for (; k < kN && i >= 1; ++k) {
            if (i <= (mask >> 1)) mask >>= 1;  // rare and predictable: i crossed a power of two
            const uint32_t v = tb[k] & mask;
            const uint32_t acc = v <= i ? 1u : 0u;
            const uint32_t j = acc ? v : i;
            const int64_t t = out[j];
            out[j] = out[i];
            out[i] = t;
            i -= acc;
        }
        pos = k;
    }
    *mt_pos = pos;
    return RFI_OK;
}
