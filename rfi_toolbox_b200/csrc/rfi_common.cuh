// rfi_common.cuh -- shared device helpers for the sm_100a hot-path kernels.
//
// Everything numeric here is written so that the float32 / float64 results equal what
// NumPy computes for the same op on the host: one IEEE operation per source-level
// operation, no FMA contraction (the build passes -fmad=false; the only fused op is the
// explicit fma inside cabs_np, which NumPy's SIMD complex-abs loop also fuses).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <type_traits>

#include "../../include/rfi_b200.h"

#define RFI_DEVINL __device__ __forceinline__

namespace rfi {

// ----------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define RFI_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return rfi::cuda_fail(_e, #expr); \
    } while (0)

// ----------------------------------------------------------------------------------------
// scalar traits: T in {float, double}; K = order-preserving unsigned key type
template <typename T> struct Scalar;
template <> struct Scalar<float> {
    using key_t = uint32_t;
    static constexpr int kBits = 32;
    RFI_DEVINL static uint32_t bits(float x) { return __float_as_uint(x); }
    RFI_DEVINL static float from_bits(uint32_t b) { return __uint_as_float(b); }
    RFI_DEVINL static float nan() { return __uint_as_float(0x7fffffffu); }
    RFI_DEVINL static float inf() { return __uint_as_float(0x7f800000u); }
    RFI_DEVINL static float sqrt_rn(float x) { return __fsqrt_rn(x); }
    RFI_DEVINL static float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    RFI_DEVINL static float fmin_nan(float a, float b) { return fminf(a, b); }
    RFI_DEVINL static float fmax_nan(float a, float b) { return fmaxf(a, b); }
    // log10 of a float32, correctly rounded in practice (fp64 evaluation, one final rounding).
    // NumPy's float32 log10 is SVML on AVX-512 hosts (<= 3 ulp off); see DESIGN.md "numerics".
    RFI_DEVINL static float log10_(float x) { return (float)::log10((double)x); }
    // phase channel only (tolerance class: NumPy's float32 arctan2 is SVML, a few ulp): CUDA's atan2f
    // (<= 2 ulp) instead of a float64 evaluation -- the complex-branch writer was bound by it
    RFI_DEVINL static float atan2_(float y, float x) { return ::atan2f(y, x); }
};
template <> struct Scalar<double> {
    using key_t = unsigned long long;
    static constexpr int kBits = 64;
    RFI_DEVINL static unsigned long long bits(double x) { return (unsigned long long)__double_as_longlong(x); }
    RFI_DEVINL static double from_bits(unsigned long long b) { return __longlong_as_double((long long)b); }
    RFI_DEVINL static double nan() { return __longlong_as_double(0x7fffffffffffffffLL); }
    RFI_DEVINL static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    RFI_DEVINL static double sqrt_rn(double x) { return __dsqrt_rn(x); }
    RFI_DEVINL static double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    RFI_DEVINL static double fmin_nan(double a, double b) { return ::fmin(a, b); }
    RFI_DEVINL static double fmax_nan(double a, double b) { return ::fmax(a, b); }
    RFI_DEVINL static double log10_(double x) { return ::log10(x); }
    RFI_DEVINL static double atan2_(double y, double x) { return ::atan2(y, x); }
};

RFI_DEVINL float fabs_(float x) { return fabsf(x); }
RFI_DEVINL double fabs_(double x) { return ::fabs(x); }
template <typename T> RFI_DEVINL bool is_nan(T x) { return x != x; }
template <typename T> RFI_DEVINL bool is_inf(T x) { return fabs_(x) == Scalar<T>::inf(); }

// Order-preserving key: ascending float order == ascending unsigned order; every NaN maps
// to the all-ones key, which doubles as the "excluded" marker (nanmedian drops NaNs).
// Branch-free: 2 logic ops + the NaN canonicalisation (3 ops); from_key needs no special case
// (the all-ones key decodes to the canonical quiet NaN).
template <typename T>
RFI_DEVINL typename Scalar<T>::key_t to_key(T x) {
    using K = typename Scalar<T>::key_t;
    using SK = typename std::make_signed<K>::type;
    constexpr int kB = Scalar<T>::kBits;
    constexpr K kSign = K(1) << (kB - 1);
    constexpr K kInfBits = sizeof(T) == 4 ? K(0x7f800000u) : (K(0x7ff00000u) << 32);
    const K b = Scalar<T>::bits(x);
    const K m = (K)((SK)b >> (kB - 1));  // all ones for negative values
    const K k = b ^ (m | kSign);
    return ((b & ~kSign) > kInfBits) ? ~K(0) : k;
}
template <typename T>
RFI_DEVINL T from_key(typename Scalar<T>::key_t k) {
    using K = typename Scalar<T>::key_t;
    using SK = typename std::make_signed<K>::type;
    constexpr int kB = Scalar<T>::kBits;
    constexpr K kSign = K(1) << (kB - 1);
    const K m = (K)((SK)k >> (kB - 1));  // all ones for keys of non-negative values
    return Scalar<T>::from_bits(k ^ (~m | kSign));
}

// |re + i*im| exactly as NumPy's SIMD complex absolute computes it (verified bit-for-bit
// against np.abs for complex64 and complex128 on AVX2/AVX-512 builds, tests/test_oracle*):
//     hi = max(|re|,|im|); lo = min(|re|,|im|); r = lo / hi; hi * sqrt(fma(r, r, 1))
// with the inf/nan rules of C hypot (inf wins over nan).
template <typename T>
RFI_DEVINL T cabs_np(T re, T im) {
    T a = fabs_(re), b = fabs_(im);
    if (is_inf(a) || is_inf(b)) return Scalar<T>::inf();
    if (is_nan(a) || is_nan(b)) return Scalar<T>::nan();
    T hi = a > b ? a : b, lo = a > b ? b : a;
    if (hi == T(0)) return T(0);
    T r = lo / hi;
    return Scalar<T>::sqrt_rn(Scalar<T>::fma(r, r, T(1))) * hi;
}

// NumPy's median of n values whose two middle order statistics are (a, b).
template <typename T>
RFI_DEVINL T median_of_pair(T a, T b, uint32_t n) {
    if (n & 1u) return a;
    T sum = a + b;
    return sum * T(0.5);
}

// keys of +inf / -inf
template <typename T>
__host__ __device__ constexpr typename Scalar<T>::key_t to_key_const_inf(bool negative) {
    using K = typename Scalar<T>::key_t;
    constexpr K kSign = K(1) << (Scalar<T>::kBits - 1);
    constexpr K kInfBits = sizeof(T) == 4 ? K(0x7f800000u) : (K(0x7ff00000u) << 32);
    return negative ? ~(kInfBits | kSign) : (kInfBits | kSign);
}

// ----------------------------------------------------------------------------------------
// block-wide reductions.  NT threads; `slots` is NT/32 words of shared scratch per buffer,
// double buffered so that back-to-back reductions need ONE __syncthreads each.
template <int NT>
struct BlockScratch {
    static constexpr int kWarps = NT / 32;
    alignas(16) unsigned long long w[2][kWarps * 2];  // two 64-bit lanes per warp (enough for min+max / 2 counts)
    int parity;
};

template <int NT>
RFI_DEVINL void scratch_init(BlockScratch<NT>& s) {
    if (threadIdx.x == 0) s.parity = 0;
}

// Sum of one u32 per thread over the block, returned to every thread.
template <int NT>
RFI_DEVINL uint32_t block_sum(uint32_t v, BlockScratch<NT>& s, int& parity) {
    constexpr int W = NT / 32;
    v = __reduce_add_sync(0xffffffffu, v);
    uint32_t* buf = reinterpret_cast<uint32_t*>(s.w[parity]);
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < W; i += 4) {
        uint4 q = *reinterpret_cast<const uint4*>(buf + i);
        t += q.x + q.y + q.z + q.w;
    }
    parity ^= 1;
    return t;
}

// Two u32 sums at once.
template <int NT>
RFI_DEVINL void block_sum2(uint32_t& a, uint32_t& b, BlockScratch<NT>& s, int& parity) {
    constexpr int W = NT / 32;
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
    uint32_t* buf = reinterpret_cast<uint32_t*>(s.w[parity]);
    if ((threadIdx.x & 31) == 0) {
        buf[threadIdx.x >> 5] = a;
        buf[W + (threadIdx.x >> 5)] = b;
    }
    __syncthreads();
    uint32_t ta = 0, tb = 0;
#pragma unroll
    for (int i = 0; i < W; i += 4) {
        uint4 q = *reinterpret_cast<const uint4*>(buf + i);
        uint4 r = *reinterpret_cast<const uint4*>(buf + W + i);
        ta += q.x + q.y + q.z + q.w;
        tb += r.x + r.y + r.z + r.w;
    }
    a = ta;
    b = tb;
    parity ^= 1;
}

RFI_DEVINL uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
RFI_DEVINL uint32_t warp_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
RFI_DEVINL unsigned long long warp_min(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u < v ? u : v;
    }
    return v;
}
RFI_DEVINL unsigned long long warp_max(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// min and max of one key per thread (keys are unsigned); result to every thread.
template <int NT, typename K>
RFI_DEVINL void block_minmax_key(K& lo, K& hi, BlockScratch<NT>& s, int& parity) {
    constexpr int W = NT / 32;
    lo = warp_min(lo);
    hi = warp_max(hi);
    unsigned long long* buf = s.w[parity];
    if ((threadIdx.x & 31) == 0) {
        buf[threadIdx.x >> 5] = (unsigned long long)lo;
        buf[W + (threadIdx.x >> 5)] = (unsigned long long)hi;
    }
    __syncthreads();
    unsigned long long a = buf[0], b = buf[W];
#pragma unroll
    for (int i = 1; i < W; ++i) {
        unsigned long long x = buf[i], y = buf[W + i];
        a = x < a ? x : a;
        b = y > b ? y : b;
    }
    lo = (K)a;
    hi = (K)b;
    parity ^= 1;
}

// NaN-ignoring min / max of one value per thread (nanmin / nanmax); NaN iff all NaN.
template <int NT, typename T>
RFI_DEVINL void block_nanminmax(T& lo, T& hi, BlockScratch<NT>& s, int& parity) {
    constexpr int W = NT / 32;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = Scalar<T>::fmin_nan(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = Scalar<T>::fmax_nan(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    double* buf = reinterpret_cast<double*>(s.w[parity]);
    if ((threadIdx.x & 31) == 0) {
        buf[threadIdx.x >> 5] = (double)lo;
        buf[W + (threadIdx.x >> 5)] = (double)hi;
    }
    __syncthreads();
    double a = buf[0], b = buf[W];
#pragma unroll
    for (int i = 1; i < W; ++i) {
        a = ::fmin(a, buf[i]);
        b = ::fmax(b, buf[W + i]);
    }
    lo = (T)a;
    hi = (T)b;
    parity ^= 1;
}

// Four NaN-ignoring (min, max) pairs per thread at once; results to every thread.  `buf` is
// NT / 32 * 8 values of shared scratch.  Warp shuffles, one store per warp, ONE barrier pair,
// then a 16-lane shuffle tree over the per-warp partials.
template <int NT, typename T>
RFI_DEVINL void block_nanminmax4(T (&lo)[4], T (&hi)[4], T* buf) {
    constexpr int W = NT / 32;
    static_assert(W <= 32, "one partial per lane");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[k] = Scalar<T>::fmin_nan(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = Scalar<T>::fmax_nan(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    __syncthreads();  // buf may alias a buffer other warps still read
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { buf[warp * 8 + k] = lo[k]; buf[warp * 8 + 4 + k] = hi[k]; }
    }
    __syncthreads();
    const int src = lane < W ? lane : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo[k] = buf[src * 8 + k]; hi[k] = buf[src * 8 + 4 + k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[k] = Scalar<T>::fmin_nan(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = Scalar<T>::fmax_nan(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    __syncthreads();  // before buf is reused
}

// ----------------------------------------------------------------------------------------
// Round counters: block-wide sum of one u32 per thread with ONE barrier and ~6 instructions
// per thread: warp REDUX, one shared atomic per warp into a rotating counter, barrier, one
// broadcast load.  Four counters rotate; counter (r+2)&3 is cleared right after round r's
// barrier (its last readers passed barrier r-1, its next writers start after barrier r+1).
struct RoundCounter {
    uint32_t c[4];
    int pad[4];
};
RFI_DEVINL void round_init(RoundCounter& rc) {
    if (threadIdx.x < 4) rc.c[threadIdx.x] = 0;
    __syncthreads();
}
RFI_DEVINL uint32_t round_sum(uint32_t v, RoundCounter& rc, int& r) {
    v = __reduce_add_sync(0xffffffffu, v);
    uint32_t* slot = rc.c + (r & 3);
    if ((threadIdx.x & 31) == 0) atomicAdd(slot, v);
    __syncthreads();
    const uint32_t t = *slot;
    if (threadIdx.x == 0) rc.c[(r + 2) & 3] = 0;
    ++r;
    return t;
}

// count += (key < trial), as ISETP (ALU pipe) + predicated FADD (FMA pipe): the two pipes
// each retire one warp instruction every other cycle per sub-partition, so this pairing
// issues at full rate where an integer add would serialise on the ALU pipe.
RFI_DEVINL void count_lt(float& c, uint32_t key, uint32_t trial) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}"
        : "+f"(c) : "r"(key), "r"(trial));
}
RFI_DEVINL void count_lt(float& c, unsigned long long key, unsigned long long trial) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u64 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}"
        : "+f"(c) : "l"(key), "l"(trial));
}
RFI_DEVINL void count_le(float& c, uint32_t key, uint32_t trial) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}"
        : "+f"(c) : "r"(key), "r"(trial));
}
RFI_DEVINL void count_le(float& c, unsigned long long key, unsigned long long trial) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.u64 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}"
        : "+f"(c) : "l"(key), "l"(trial));
}

// ----------------------------------------------------------------------------------------
// Register-resident exact selection.
//
// Each of the NT threads holds E keys; excluded samples carry the all-ones key and are not
// counted in n.  block_select2 returns the keys of 0-based ranks k1 and k2 (k2 = k1 or k1+1:
// the two middle order statistics of NumPy's median).
//
// Method: MSB-first binary radix select ("bit bisection").  Round b asks how many keys are
// below prefix|1<<b; a round costs 2 instructions per key and one block reduction, no
// shared-memory histogram (ATOMS retires ~0.5 key/clk/SM on Blackwell, ~20x slower than
// this).  The leading bits shared by min and max are skipped.
// Candidate compaction: once at most kSelCap keys still share the decided prefix they are
// copied to a shared list and the remaining bits are resolved two per round on <= kSelCap/NT
// keys per thread (three trial keys per round, their counts packed 10 bits each into one
// word), instead of sweeping all E keys per thread for every remaining bit.
constexpr int kSelCap = 1023;  // counts must fit 10 bits

template <typename K>
struct SelectScratch {
    K list[kSelCap + 1];
    uint32_t cursor;
};

template <int NT, int E, typename K>
RFI_DEVINL void block_select2(const K (&key)[E], uint32_t n, uint32_t k1, uint32_t k2,
                              K& out1, K& out2, BlockScratch<NT>& s, int& parity,
                              RoundCounter& rc, int& round, SelectScratch<K>& ss) {
    constexpr K kExcl = ~K(0);
    constexpr int kBits = sizeof(K) * 8;
    constexpr int CE = (kSelCap + NT) / NT;  // compacted keys per thread
    if (n == 0) {  // uniform: n comes from a block reduction
        out1 = out2 = kExcl;
        return;
    }
    K lo = kExcl, hi = 0;
    uint32_t lt_upper = 0;  // this thread's keys below prefix + 2^(b+1) (initially: its valid keys)
#pragma unroll
    for (int e = 0; e < E; ++e) {
        K x = key[e];
        lo = x < lo ? x : lo;
        const bool ex = (x == kExcl);
        K y = ex ? K(0) : x;
        hi = y > hi ? y : hi;
        lt_upper += ex ? 0u : 1u;
    }
    block_minmax_key<NT, K>(lo, hi, s, parity);
    K prefix = lo;
    const K diff = lo ^ hi;
    bool compacted = false;
    uint32_t below = 0, upper = n;  // keys < prefix ; keys < prefix + 2^(b+1)
    uint32_t lt_below = 0;          // this thread's keys below prefix
    K ck[CE];
    if (diff != 0) {
        const int hb = (kBits - 1) - (sizeof(K) == 8 ? __clzll((long long)diff) : __clz((int)diff));
        prefix = (hb == kBits - 1) ? K(0) : (lo >> (hb + 1)) << (hb + 1);
        int b = hb;
#pragma unroll 1
        for (; b >= 0; --b) {
            if (upper - below <= (uint32_t)kSelCap && b >= 1) break;  // uniform
            const K trial = prefix | (K(1) << b);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
            for (int e = 0; e < E; e += 4) {
                count_lt(c0, key[e], trial);
                count_lt(c1, key[e + 1], trial);
                count_lt(c2, key[e + 2], trial);
                count_lt(c3, key[e + 3], trial);
            }
            const uint32_t lt = (uint32_t)((c0 + c1) + (c2 + c3));
            const uint32_t c = round_sum(lt, rc, round);
            if (c <= k1) { prefix = trial; below = c; lt_below = lt; } else { upper = c; lt_upper = lt; }
        }
        if (b >= 0) {
            // ---- compact the keys whose bits above b equal the prefix
            compacted = true;
            if (threadIdx.x == 0) ss.cursor = 0;
            __syncthreads();
            // this thread's candidates = its keys in [prefix, prefix + 2^(b+1)), already counted
            // by the rounds that last moved the two bounds -> no counting pass
            const uint32_t mine = lt_upper - lt_below;
            uint32_t at = mine ? atomicAdd(&ss.cursor, mine) : 0u;
            const int sh = b + 1;  // undecided low bits
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const K x = key[e] ^ prefix;
                const bool is = (sh >= kBits || (x >> sh) == 0) && key[e] != kExcl;
                if (is) ss.list[at++] = key[e];
            }
            __syncthreads();
            const uint32_t ncand = upper - below;
#pragma unroll
            for (int j = 0; j < CE; ++j) {
                const uint32_t idx = threadIdx.x + j * NT;
                ck[j] = idx < ncand ? ss.list[idx] : kExcl;
            }
            const uint32_t kk = k1 - below;  // rank inside the candidate list
            // ---- two bits per round
#pragma unroll 1
            for (; b >= 1; b -= 2) {
                const K t1 = prefix | (K(1) << (b - 1)), t2 = prefix | (K(2) << (b - 1)),
                        t3 = prefix | (K(3) << (b - 1));
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < CE; ++j)
                    c += (ck[j] < t1 ? 1u : 0u) + (ck[j] < t2 ? 1u << 10 : 0u) + (ck[j] < t3 ? 1u << 20 : 0u);
                c = round_sum(c, rc, round);
                const uint32_t n1 = c & 1023u, n2 = (c >> 10) & 1023u, n3 = c >> 20;
                prefix = (n3 <= kk) ? t3 : (n2 <= kk) ? t2 : (n1 <= kk) ? t1 : prefix;
            }
            if (b == 0) {  // odd number of remaining bits: one single-bit round
                const K t1 = prefix | K(1);
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < CE; ++j) c += (ck[j] < t1) ? 1u : 0u;
                c = round_sum(c, rc, round);
                if (c <= kk) prefix = t1;
            }
        }
    }
    out1 = prefix;
    if (k2 == k1) {
        out2 = prefix;
        return;
    }
    // rank k1+1: same key if duplicates reach it, else the smallest key above.
    if (compacted) {
        uint32_t cle = 0;
        K nxt = kExcl;
#pragma unroll
        for (int j = 0; j < CE; ++j) {
            cle += (ck[j] <= prefix) ? 1u : 0u;
            const K y = (ck[j] > prefix) ? ck[j] : kExcl;
            nxt = y < nxt ? y : nxt;
        }
        cle = round_sum(cle, rc, round);
        if (k2 < below + cle) {  // uniform
            out2 = prefix;
            return;
        }
        K dummy = 0;
        block_minmax_key<NT, K>(nxt, dummy, s, parity);
        if (nxt != kExcl) {  // a candidate above: smaller than every key outside the list
            out2 = nxt;
            return;
        }
        // the answer is the largest candidate: the next key lies outside the list (below)
    } else {
        float cle = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) count_le(cle, key[e], prefix);
        const uint32_t ncle = round_sum((uint32_t)cle, rc, round);
        if (k2 < ncle) {  // uniform
            out2 = prefix;
            return;
        }
    }
    K nxt = kExcl;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const K y = (key[e] > prefix) ? key[e] : kExcl;
        nxt = y < nxt ? y : nxt;
    }
    K dummy = 0;
    block_minmax_key<NT, K>(nxt, dummy, s, parity);
    out2 = nxt;
}

// NumPy median of the n valid keys held by the block: mean of the two middle order
// statistics in T ((a + b) / 2 with one rounding of the sum), NaN when n == 0.
// The two order statistics themselves are returned through lo_key / hi_key.
template <typename T, int NT, int E>
RFI_DEVINL T block_median(const typename Scalar<T>::key_t (&key)[E], uint32_t n,
                          BlockScratch<NT>& s, int& parity, RoundCounter& rc, int& round,
                          SelectScratch<typename Scalar<T>::key_t>& ss,
                          typename Scalar<T>::key_t* lo_key = nullptr,
                          typename Scalar<T>::key_t* hi_key = nullptr) {
    using K = typename Scalar<T>::key_t;
    if (n == 0) {
        if (lo_key) *lo_key = ~K(0);
        if (hi_key) *hi_key = ~K(0);
        return Scalar<T>::nan();
    }
    K a, b;
    block_select2<NT, E, K>(key, n, (n - 1) >> 1, n >> 1, a, b, s, parity, rc, round, ss);
    if (lo_key) *lo_key = a;
    if (hi_key) *hi_key = b;
    return median_of_pair<T>(from_key<T>(a), from_key<T>(b), n);
}

}  // namespace rfi
