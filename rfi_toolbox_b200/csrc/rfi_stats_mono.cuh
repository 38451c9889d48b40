// rfi_stats_mono.cuh -- phase 1 of create_dataset for the common tile: every sample >= +0 and
// finite after the stretch (magnitudes; no exact zero under LOG10).
//
// For such a tile every stage of the reference chain -- x / m (m > 0), sqrt, log10, / m2 -- is
// monotone non-decreasing after rounding, so the whole job can be done on the RAW samples:
//
//   * median: the normalisation median is an order statistic of the raw tile, and the flag
//     centre is the image of the same two middle order statistics;
//   * MAD: d(a) = |proc(a) - c| is V-shaped in the raw value a, so "the r smallest deviations"
//     is a contiguous window of the raw order around the centre;
//   * labels: proc(a) > thr_hi  <=>  a > raw_hi, with raw_hi = max{a : proc(a) <= thr_hi} found
//     exactly by a 32-ary search that evaluates proc() itself -- phase 2 then labels a pixel
//     with two compares on the raw sample, no division or square root.
//
// Selection is Floyd-Rivest style instead of bit-by-bit: a stratified 512-sample (every row and
// every column of the tile contributes 4 samples) is sorted (four warps sort 128 keys each in
// registers, then every thread ranks one key in the other three runs); two sample ranks 3 sigma
// either side of the target bracket it; one pass over the 16384 keys counts what lies below the
// bracket and marks the ~2300 keys inside it (one bit per key), the marked keys are then gathered
// into a list while the 512-bucket linear histogram over the bracket is filled; the exact rank inside
// that list is resolved from that histogram + one <= 64 element ranking.  All brackets are validated (the
// answer must fall strictly inside what was proven), otherwise the tile is handed to the general
// kernel (tile_stats_general, RFI_TILE_GENERAL), which then runs in the same CTA -- results are
// exact either way.
//
// Cost: ~64 k warp instructions per tile instead of ~150 k for the 32-round register bisection
// (a fifth of them the load with the NumPy-exact |z|; DESIGN.md 5.1 lists where the rest goes).  float32 keys: half in shared memory, half in a
// thread-private global scratch (57 KB of shared memory per CTA, 3 CTAs / SM; see the kernel);
// float64 keys: all in shared memory, 1 CTA / SM.
#pragma once
#include <type_traits>
#include "rfi_tiles.cuh"

namespace rfi {

constexpr int kMonoNT = 512;     // threads = sample size
constexpr int kMonoCap = 4096;   // candidate list capacity (expected ~2300 at 3 sigma)
constexpr int kMonoBuckets = 512;
constexpr int kMonoSmall = 64;   // max elements of the bucket that holds the answer
#ifndef RFI_MONO_SIGMA
#define RFI_MONO_SIGMA 3.0f
#endif
// half width of the sample-rank brackets, in sigma of a sample rank.  Measured on the bench workload:
// 3.0 -> 1.198 ms, 3.5 -> 1.247 ms, 4.0 -> 1.294 ms (every extra candidate costs more than the
// ~0.9 % of tiles whose bracket misses and that fall back to the general algorithm).
constexpr float kMonoSigma = RFI_MONO_SIGMA;

template <typename K>
struct MonoShared {          // static part (must stay small: 3 CTAs / SM)
    uint32_t hist[kMonoBuckets];
    K small_a[kMonoSmall], small_b[kMonoSmall];
    uint32_t cursor, below, n_small_a, n_small_b;
    uint32_t acc[4];
    K kmin, kmax;
    K res1, res2;
    uint32_t b1, b2, pre1, pre2, cnt1, cnt2;
    int win[4];
    int fail;
};

// ---- one warp sorts 512 keys held 16 per lane, element index e = r * 32 + lane ------------
template <typename K>
RFI_DEVINL K shfl_xor_key(K v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// one warp sorts 32 * R keys held R per lane, element index e = r * 32 + lane (ascending)
template <typename K, int R>
RFI_DEVINL void warp_sort_regs(K (&v)[R], int lane) {
    constexpr int N = 32 * R;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        // in-lane stages: partner register r ^ (j / 32); direction fixed by r at compile time
#pragma unroll
        for (int j = k >> 1; j >= 32; j >>= 1) {
            const int jr = j >> 5;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((r & jr) == 0) {
                    const int r2 = r | jr;
                    const bool up = (((r << 5) & k) == 0);
                    const K a = v[r], b = v[r2];
                    const K mn = a < b ? a : b, mx = a < b ? b : a;
                    v[r] = up ? mn : mx;
                    v[r2] = up ? mx : mn;
                }
            }
        }
        // cross-lane stages
        const int jstart = (k >> 1) < 16 ? (k >> 1) : 16;
#pragma unroll
        for (int j = jstart; j > 0; j >>= 1) {
            const bool lower = (lane & j) == 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool up = ((((r << 5) | lane) & k) == 0);
                const K o = shfl_xor_key<K>(v[r], j);
                const bool keep_min = (lower == up);
                const K mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
                v[r] = keep_min ? mn : mx;
            }
        }
    }
}
template <typename K>
RFI_DEVINL void warp_sort512(K (&v)[16], int lane) { warp_sort_regs<K, 16>(v, lane); }

// raw key of a non-negative sample: its bit pattern (orders like the value); NaN -> all ones
template <typename T>
RFI_DEVINL typename Scalar<T>::key_t raw_key(T a) {
    using K = typename Scalar<T>::key_t;
    return is_nan(a) ? ~K(0) : Scalar<T>::bits(a);
}
template <typename T>
RFI_DEVINL T raw_val(typename Scalar<T>::key_t k) { return Scalar<T>::from_bits(k); }

// processed sample WITHOUT the inf fill (the searches walk outside the tile's value range).
// `mode` packs the plan's switches (kProc*) so that the out-of-line chain reads no plan field.
constexpr int kProcDivM = 1, kProcSqrt = 2, kProcLog10 = 4, kProcDivM2 = 8;
template <typename T>
RFI_DEVINL int proc_mode_of(const PlanDev& p, T m, T m2) {
    return ((p.norm_before && m > T(0)) ? kProcDivM : 0) | (p.stretch == RFI_STRETCH_SQRT ? kProcSqrt : 0) |
           (p.stretch == RFI_STRETCH_LOG10 ? kProcLog10 : 0) | ((p.norm_after && m2 > T(0)) ? kProcDivM2 : 0);
}
// The optional steps are real branches (pin() keeps the compiler from evaluating a division that
// is not asked for and selecting afterwards: with its divisor 0 -- a median that is not in use --
// such a division takes div.rn's out-of-line path, ~45 instructions per call for nothing).
RFI_DEVINL void pin(float& a) { asm volatile("" : "+f"(a)); }
RFI_DEVINL void pin(double& a) { asm volatile("" : "+d"(a)); }
template <typename T>
__device__ __noinline__ T proc_mode(T a, int mode, T m, T m2) {
    if (mode & kProcDivM) { pin(a); a = a / m; }
    if (mode & kProcSqrt) a = Scalar<T>::sqrt_rn(fabs_(a));
    else if (mode & kProcLog10) a = Scalar<T>::log10_(fabs_(a));
    if (mode & kProcDivM2) { pin(a); a = a / m2; }
    return a;
}
template <typename T>
RFI_DEVINL T proc_nofill(T a, const PlanDev& p, T m, T m2) { return proc_mode<T>(a, proc_mode_of<T>(p, m, m2), m, m2); }

// Exact rank resolution inside the candidate list cand[0 .. M): keys of ranks q1 <= q2 <= q1 + 1.
// Block-wide, uniform control flow.  Rank q1 by a linear 512-bucket histogram over the list's
// key range, refined on the answer's bucket until that bucket holds <= 64 keys (ranked by one
// warp) or a single key value (duplicates); rank q2 = q1 + 1 is then either the same key or the
// smallest key above it.
// bucket shift of the <= 512-bucket histogram over a key range of the given width
template <typename K>
RFI_DEVINL int mono_bucket_shift(K width) {
    constexpr int kBitsK = (int)sizeof(K) * 8;
    const int wbits = width == 0 ? 0 : kBitsK - (sizeof(K) == 8 ? __clzll((long long)width) : __clz((int)width));
    return wbits > 9 ? wbits - 9 : 0;
}
// `prebuilt`: the caller knows a range [klo_in, khi_in] that holds every key of the list and has ALREADY
// filled sh.hist over it (bucket (x - klo_in) >> mono_bucket_shift(khi_in - klo_in)) while it wrote the list,
// behind a barrier -- the list is then neither scanned for its range nor for the first histogram.
template <typename K, int NT>
RFI_DEVINL void mono_resolve(const K* cand, uint32_t M, uint32_t q1, uint32_t q2, K& out1, K& out2,
                             MonoShared<K>& sh, bool prebuilt = false, K klo_in = 0, K khi_in = 0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    K klo = klo_in, khi = khi_in;
    if (!prebuilt) {
        // ---- range of the list
        K lo = ~K(0), hi = 0;
        for (uint32_t i = tid; i < M; i += NT) {
            const K x = cand[i];
            lo = x < lo ? x : lo;
            hi = x > hi ? x : hi;
        }
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (tid == 0) { sh.kmin = ~K(0); sh.kmax = 0; }
        __syncthreads();
        if (lane == 0) {
            if (sizeof(K) == 8) {
                atomicMin(reinterpret_cast<unsigned long long*>(&sh.kmin), (unsigned long long)lo);
                atomicMax(reinterpret_cast<unsigned long long*>(&sh.kmax), (unsigned long long)hi);
            } else {
                atomicMin(reinterpret_cast<unsigned int*>(&sh.kmin), (unsigned int)lo);
                atomicMax(reinterpret_cast<unsigned int*>(&sh.kmax), (unsigned int)hi);
            }
        }
        __syncthreads();
        klo = sh.kmin; khi = sh.kmax;
    }
    uint32_t q = q1;  // rank inside [klo, khi]
    K answer = klo;
    // rank q1 + 1 is usually found on the way: when the answer's bucket is ranked directly and the answer is not
    // the bucket's last key, the key ranked right behind it is the next order statistic
    bool have_next = false;
    K next_key = 0;
    for (int iter = 0; iter < 8; ++iter) {  // <= ceil(64 / 9) refinements
        const K width = khi - klo;
        const int shf = mono_bucket_shift<K>(width);  // <= 512 buckets
        if (prebuilt && iter == 0) {
            if (tid == 0) { sh.n_small_a = 0; sh.n_small_b = 0; }
        } else {
            for (int b = tid; b < kMonoBuckets; b += NT) sh.hist[b] = 0;
            if (tid == 0) { sh.n_small_a = 0; sh.n_small_b = 0; }
            __syncthreads();
            for (uint32_t i = tid; i < M; i += NT) {
                const K x = cand[i];
                if ((K)(x - klo) <= width) atomicAdd(&sh.hist[(uint32_t)((x - klo) >> shf)], 1u);
            }
        }
        __syncthreads();
        if (warp == 0) {  // bucket of rank q
            uint32_t local[16], sum = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { local[i] = sh.hist[lane * 16 + i]; sum += local[i]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t pre = incl - sum;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t c = local[i];
                if (q >= pre && q < pre + c) { sh.b1 = lane * 16 + i; sh.pre1 = pre; sh.cnt1 = c; }
                pre += c;
            }
        }
        __syncthreads();
        const uint32_t b = sh.b1, pre = sh.pre1, cnt = sh.cnt1;
        const K blo = klo + ((K)b << shf);
        if (shf == 0) { answer = blo; break; }  // one key value per bucket
        const K bw = (K(1) << shf) - 1;
        if (cnt <= (uint32_t)kMonoSmall) {
            for (uint32_t i = tid; i < M; i += NT) {
                const K x = cand[i];
                if ((K)(x - blo) <= bw) sh.small_a[atomicAdd(&sh.n_small_a, 1u)] = x;
            }
            __syncthreads();
            {   // element with exactly (q - pre) smaller-or-earlier elements: 8 threads per element
                const uint32_t n = sh.n_small_a, want = q - pre;
                const uint32_t i = tid >> 3, sub = tid & 7;
                const K x = i < n ? sh.small_a[i] : ~K(0);
                uint32_t rank = 0;
                for (uint32_t j = sub; j < n; j += 8) {
                    const K y = sh.small_a[j];
                    rank += (y < x || (y == x && j < i)) ? 1u : 0u;
                }
                rank += __shfl_xor_sync(0xffffffffu, rank, 1);
                rank += __shfl_xor_sync(0xffffffffu, rank, 2);
                rank += __shfl_xor_sync(0xffffffffu, rank, 4);
                if (sub == 0 && i < n && rank == want) sh.res1 = x;
                if (sub == 0 && i < n && rank == want + 1) { sh.small_b[0] = x; sh.n_small_b = 1; }
            }
            __syncthreads();
            answer = sh.res1;
            have_next = sh.n_small_b != 0;
            next_key = sh.small_b[0];
            break;
        }
        klo = blo; khi = blo + bw; q -= pre;  // refine on the answer's bucket
        __syncthreads();
    }
    out1 = answer;
    out2 = answer;
    if (q2 != q1 && have_next) {
        out2 = next_key;
    } else if (q2 != q1) {  // rank q1 + 1: the same key if duplicates reach it, else the smallest key above
        uint32_t cle = 0;
        K nxt = ~K(0);
        for (uint32_t i = tid; i < M; i += NT) {
            const K x = cand[i];
            cle += (x <= answer) ? 1u : 0u;
            const K y = x > answer ? x : ~K(0);
            nxt = y < nxt ? y : nxt;
        }
        cle = __reduce_add_sync(0xffffffffu, cle);
        nxt = warp_min(nxt);
        if (tid == 0) { sh.cnt2 = 0; sh.kmin = ~K(0); }
        __syncthreads();
        if (lane == 0) {
            atomicAdd(&sh.cnt2, cle);
            if (sizeof(K) == 8) atomicMin(reinterpret_cast<unsigned long long*>(&sh.kmin), (unsigned long long)nxt);
            else atomicMin(reinterpret_cast<unsigned int*>(&sh.kmin), (unsigned int)nxt);
        }
        __syncthreads();
        if (q2 >= sh.cnt2) out2 = sh.kmin;
    }
    __syncthreads();
}

// general algorithm (rfi_tiles.cu): any tile, register-resident radix select
template <int DT, int NT>
__device__ __noinline__ void tile_stats_general(const PlanDev& p, const void* __restrict__ data,
                                   const uint8_t* __restrict__ flags, rfi_tile_stat_t* __restrict__ stats,
                                   int route_bits, typename Scalar<typename In<DT>::T>::key_t* stash_override = nullptr);

// GK (float32 keys): only the first kMonoGS of the 8 key groups stay in shared memory; the rest
// live in a global scratch (64 KB per tile, of which the upper part is used here and all of it by
// the general fallback).  The scratch is THREAD-PRIVATE -- a thread only ever re-reads the keys it
// wrote itself, with coalesced 16-byte accesses -- so it needs no fences and is served by L1 / L2.
// Shared memory per CTA drops from 88 KB to 57 KB and three CTAs fit one SM instead of two:
// more independent barrier domains per scheduler for a kernel whose warps mostly wait at
// barriers (measured: 1.405 -> 1.297 ms on the bench workload; 1.36 ms with all keys in the
// scratch, 1.32 / 1.335 ms with 6 / 5 groups in shared memory).
constexpr int kMonoGKB = 3, kMonoGS = 4;

// The body of phase 1 as a device function, shared by the stand-alone statistics kernel below and by
// the single-launch kernel of rfi_tiles.cu (tile_fused_kernel, FUSED = true), which goes on to write
// the tile's patches from the magnitudes this function leaves in shared memory.  FUSED changes two
// things: (1) the key tile uses the WRITER's layout -- element (row, col) at row * 128 +
// (col ^ (row & 31)), conflict-free by rows and by columns; aligned groups of four columns stay
// aligned groups, so every access below is still one 128-bit access per thread, only the slot of the
// group and the order inside it change; (2) the tile's statistics also go to `st_sh` (shared memory).
// Returns true when the monotone algorithm finished (keys intact, raw thresholds valid), false when
// the tile was handed to the general algorithm (shared memory clobbered, statistics in `stats[tile]`).
template <int DT, int NT, bool GK, int GS, bool FUSED>
RFI_DEVINL bool mono_tile_stats(const PlanDev& p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                                rfi_tile_stat_t* __restrict__ stats,
                                typename Scalar<typename In<DT>::T>::key_t* __restrict__ gkeys,
                                unsigned char* smem_raw, MonoShared<typename Scalar<typename In<DT>::T>::key_t>& sh,
                                rfi_tile_stat_t* st_sh) {
    using T = typename In<DT>::T;
    using K = typename Scalar<T>::key_t;
    static_assert(NT == kMonoNT, "one sample per thread");
    static_assert(!(FUSED && GK), "the fused kernel keeps the whole tile in shared memory");
    constexpr int E = kP * kP / NT;  // 32 keys per thread
    constexpr int G = E / 4;
    constexpr int RS = NT / 32;
    constexpr K kExcl = ~K(0);
    constexpr K kInfKey = sizeof(T) == 4 ? K(0x7f800000u) : (K(0x7ff00000u) << 32);
    constexpr K kSignBit = K(1) << (Scalar<T>::kBits - 1);

    // keys: [E/4][NT][4]; with GK the first GS groups stay in shared memory, the rest live in the
    // thread-private global scratch
    K* skeys = reinterpret_cast<K*>(smem_raw);
    K* gk = GK ? gkeys + (size_t)blockIdx.x * (kP * kP) : nullptr;
    K* cand = GK ? skeys + (size_t)GS * NT * 4 : skeys + kP * kP;                             // [kMonoCap]
    K* samp = cand + kMonoCap;                  // [NT] sorted raw sample
    const int tid_ = threadIdx.x;
    auto kp = [&](int g) -> K* {   // this thread's 4 keys of group g
        if (GK && g >= GS) return gk + ((size_t)g * NT + tid_) * 4;
        if (FUSED) return skeys + ((size_t)g * NT + (tid_ ^ (((g * RS + (tid_ >> 5)) >> 2) & 7))) * 4;
        return skeys + ((size_t)g * NT + tid_) * 4;
    };

    // the thread's keys whose bit is set in `marked` (bit e = element g * 4 + i), appended to cand[at ...] in
    // ascending element order -- what a sweep over all 32 keys would append, at the cost of the marked ones
    // only (one loop per address space, so that neither needs generic addressing)
    // With HIST the keys are also counted into sh.hist (bucket (key - hlo) >> hshf): mono_resolve's first
    // histogram, built on the way.
    auto gather_marked = [&](uint32_t marked, uint32_t at, auto hist_tag, K hlo, int hshf) {
        constexpr bool HIST = decltype(hist_tag)::value;
        uint32_t ms = (GK && GS < G) ? (marked & ((1u << (4 * (GS < G ? GS : 0))) - 1u)) : marked;
        while (ms) {
            const int e = __ffs((int)ms) - 1;
            ms &= ms - 1u;
            const int g = e >> 2;
            const K* q = FUSED ? skeys + ((size_t)g * NT + (tid_ ^ (((g * RS + (tid_ >> 5)) >> 2) & 7))) * 4
                               : skeys + ((size_t)g * NT + tid_) * 4;
            const K x = q[e & 3];
            cand[at++] = x;
            if constexpr (HIST) atomicAdd(&sh.hist[(uint32_t)((K)(x - hlo) >> hshf)], 1u);
        }
        if constexpr (GK && GS < G) {
            uint32_t mg = marked >> (4 * GS);
            const K* mine_g = gk + ((size_t)GS * NT + tid_) * 4;
            while (mg) {
                const int e = __ffs((int)mg) - 1;
                mg &= mg - 1u;
                const K x = mine_g[(size_t)(e >> 2) * NT * 4 + (e & 3)];
                cand[at++] = x;
                if constexpr (HIST) atomicAdd(&sh.hist[(uint32_t)((K)(x - hlo) >> hshf)], 1u);
            }
        }
    };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile = blockIdx.x;
    const int per = p.nh * p.nw;
    const long long w = tile / per;
    const int ti = (int)((tile % per) / p.nw), tj = (int)(tile % p.nw);
    const size_t origin = ((size_t)w * p.channels + (size_t)ti * kP) * p.times + (size_t)tj * kP;

    // uniform early-out helper: hand the tile to the general kernel
    // (the reason is kept in the upper bits of `route` until the general kernel overwrites it;
    //  scripts/diag_stats.py reads it)
    auto give_up = [&](int reason) {
        __syncthreads();  // shared memory is handed over
        tile_stats_general<DT, NT>(p, data, flags, stats, RFI_TILE_GENERAL | (reason << 8), GK ? gk : nullptr);
        return false;
    };

    // ---- load: magnitude fused into the 128-bit loads, raw bit patterns to shared memory.
    // Special values cost ONE instruction per sample here: the running unsigned max of the bit
    // patterns is >= the +inf pattern iff the tile holds a NaN, an inf or a negative value
    // (sign bit); only then does a second pass classify them.
    if (tid < 4) sh.acc[tid] = 0;
    if (tid == 0) { sh.cursor = 0; sh.below = 0; sh.kmin = kExcl; sh.kmax = 0; sh.fail = 0; }
    __syncthreads();
    // this thread's sample, element e_s = g * 4 + i of its 32 (row g * 16 + warp, column
    // 4 * lane + i): stratified so that every row AND every column of the tile gives 4 samples
    const int e_s = (((lane + warp * 3) & 7) << 2) | (((lane >> 3) + warp) & 3);
    K bmax = 0, bmin = kExcl;
    // (real input under the 40-register cap of the 3-CTA variant: the fully unrolled loop hoists all
    //  32 loads and spills; two groups in flight are enough for a kernel that uses 1 TB/s of HBM)
#pragma unroll (GK && !In<DT>::cplx ? 2 : G)
    for (int g = 0; g < G; ++g) {
        const size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
        T q[4];
        load4_mag_fast<DT>(data, idx, q);
        K k4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            k4[i] = Scalar<T>::bits(q[i]);
            bmax = k4[i] > bmax ? k4[i] : bmax;
            bmin = k4[i] < bmin ? k4[i] : bmin;
        }
        if (sizeof(K) == 4) {
            if constexpr (FUSED) {
                // writer layout: column c of row r sits at c ^ (r & 31); r & 3 == warp & 3 here
                K t;
                if (warp & 1) { t = k4[0]; k4[0] = k4[1]; k4[1] = t; t = k4[2]; k4[2] = k4[3]; k4[3] = t; }
                if (warp & 2) { t = k4[0]; k4[0] = k4[2]; k4[2] = t; t = k4[1]; k4[1] = k4[3]; k4[3] = t; }
            }
            const uint4 packed = make_uint4((uint32_t)k4[0], (uint32_t)k4[1], (uint32_t)k4[2], (uint32_t)k4[3]);
            *reinterpret_cast<uint4*>(kp(g)) = packed;
            // complex input: the scratch tile ends up holding ALL magnitudes, row-major (element
            // (g * 16 + warp, 4 * lane + i) sits at index row * 128 + col), and phase 2 reads them
            // back (4 B / px) instead of the complex samples (8 B / px + |z| again)
            if constexpr (GK && In<DT>::cplx) {
                if (g < GS) *reinterpret_cast<uint4*>(gk + ((size_t)g * NT + tid_) * 4) = packed;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) kp(g)[i] = k4[i];
        }
    }
    bmax = warp_max(bmax);
    bmin = warp_min(bmin);
    if (lane == 0) {
        if (sizeof(K) == 8) {
            atomicMin(reinterpret_cast<unsigned long long*>(&sh.kmin), (unsigned long long)bmin);
            atomicMax(reinterpret_cast<unsigned long long*>(&sh.kmax), (unsigned long long)bmax);
        } else {
            atomicMin(reinterpret_cast<unsigned int*>(&sh.kmin), (unsigned int)bmin);
            atomicMax(reinterpret_cast<unsigned int*>(&sh.kmax), (unsigned int)bmax);
        }
    }
    __syncthreads();
    K tile_min = sh.kmin, tile_max = sh.kmax;  // raw extremes (valid samples)
    uint32_t nv = kP * kP;
    if (tile_max >= kInfKey) {
        // rare: NaN (-> excluded key), +inf or a negative value (-> general kernel)
        uint32_t nnan = 0, nodd = 0;
        K vmax = 0, vmin = kExcl;
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                K& x = kp(g)[i];
                const K b = x;
                const bool nan = (b & ~kSignBit) > kInfKey;
                nodd += (!nan && ((b & kSignBit) != 0 || b == kInfKey)) ? 1u : 0u;
                nnan += nan ? 1u : 0u;
                if (nan) x = kExcl;
                else { vmax = b > vmax ? b : vmax; vmin = b < vmin ? b : vmin; }
            }
        }
        __syncthreads();  // every thread has read the first-pass extremes
        if (tid == 0) { sh.kmin = kExcl; sh.kmax = 0; }
        __syncthreads();
        nnan = __reduce_add_sync(0xffffffffu, nnan);
        nodd = __reduce_add_sync(0xffffffffu, nodd);
        vmax = warp_max(vmax);
        vmin = warp_min(vmin);
        if (lane == 0) {
            if (nnan) atomicAdd(&sh.acc[0], nnan);
            if (nodd) atomicAdd(&sh.acc[1], nodd);
            if (sizeof(K) == 8) {
                atomicMin(reinterpret_cast<unsigned long long*>(&sh.kmin), (unsigned long long)vmin);
                atomicMax(reinterpret_cast<unsigned long long*>(&sh.kmax), (unsigned long long)vmax);
            } else {
                atomicMin(reinterpret_cast<unsigned int*>(&sh.kmin), (unsigned int)vmin);
                atomicMax(reinterpret_cast<unsigned int*>(&sh.kmax), (unsigned int)vmax);
            }
        }
        __syncthreads();
        nv -= sh.acc[0];
        tile_min = sh.kmin; tile_max = sh.kmax;
        if (sh.acc[1] != 0) return give_up(1);
    }
    if (nv < 64) return give_up(1);

    // ---- sort the 512-sample: warps 0..3 each sort 128 samples in registers (bitonic network, 4
    //      keys per lane, shuffles across lanes), then EVERY thread finds the global rank of one
    //      sorted key by binary searches in the other three runs (ties broken by run index, so the
    //      ranks are a permutation) and scatters it.  (16 runs of 32 cost 15 searches per key.)
    {
        K x = 0;  // own store: no barrier needed
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (g == (e_s >> 2)) x = kp(g)[e_s & 3];
        K* runs = cand;  // [4][128], free until the first compaction
        runs[tid] = x;
        const int nvalid_s = __syncthreads_count(x != kExcl);
        if (warp < 4) {
            K v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = runs[warp * 128 + r * 32 + lane];
            warp_sort_regs<K, 4>(v, lane);
#pragma unroll
            for (int r = 0; r < 4; ++r) runs[warp * 128 + r * 32 + lane] = v[r];
        }
        __syncthreads();
        const int run_id = tid >> 7;
        x = runs[tid];
        uint32_t rank = tid & 127;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (o == run_id) continue;  // warp-uniform
            const K* run = runs + o * 128;
            const bool incl = o < run_id;  // earlier runs win ties
            uint32_t pos = 0;
#pragma unroll
            for (int step = 64; step > 0; step >>= 1) {
                const K y = run[pos + step - 1];
                pos += (incl ? (y <= x) : (y < x)) ? step : 0;
            }
            const K y = run[127];
            pos += (pos == 127 && (incl ? (y <= x) : (y < x))) ? 1u : 0u;
            rank += pos;
        }
        samp[rank] = x;
        if (tid == 0) sh.acc[2] = (uint32_t)nvalid_s;
    }
    __syncthreads();
    const int sv = (int)sh.acc[2];  // valid samples (sorted first)
    if (sv < 64) return give_up(2);
    const int delta = (int)(kMonoSigma * 0.5f * sqrtf((float)sv)) + 2;  // kMonoSigma sigma of a sample rank

    const bool real_branch = !In<DT>::cplx || p.magnitude;
    const bool need_median = real_branch && (p.norm_before || p.norm_after) || p.flag_mode == RFI_FLAGS_MAD;
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;

    // ================= median of the raw tile (two middle order statistics) =================
    K v1k = 0, v2k = 0;
    if (need_median) {
        // A bracket that misses the target rank (~0.9 % of tiles at 3 sigma) is retried ONCE, on the side
        // the counts point to (the adjoining 6 sigma of sample ranks: one more counting sweep, the same
        // number of candidates), before the tile goes to the general algorithm (~8 monotone tiles).
        const int rho = (int)(((float)(2 * k1 + 1) * (float)sv) / (float)(2 * nv));
        const int ilo = rho - delta, ihi = rho + delta + 1;
        K lo = ilo < 0 ? K(0) : samp[ilo];
        K hi = ihi >= sv ? kExcl - 1 : samp[ihi];
        uint32_t M = 0, B = 0;
#pragma unroll 1
        for (int attempt = 0;; ++attempt) {
            const K span = hi - lo;
            for (int b = tid; b < kMonoBuckets; b += NT) sh.hist[b] = 0;   // filled by the compaction (barrier below)
            // (the thread's keys inside the bracket are remembered as one bit each: the compaction below
            //  visits only those -- about one key in seven -- instead of sweeping the tile again)
            // (counts as population counts of bit masks: a select and half a three-input add per key and mask)
            uint32_t lowmask = 0, inmask = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                K k4[4];
                if (sizeof(K) == 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(kp(g));
                    k4[0] = q.x; k4[1] = q.y; k4[2] = q.z; k4[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) k4[i] = kp(g)[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    lowmask += (k4[i] < lo) ? (1u << (g * 4 + i)) : 0u;
                    inmask += ((K)(k4[i] - lo) <= span) ? (1u << (g * 4 + i)) : 0u;
                }
            }
            const uint32_t below = __popc(lowmask), mine = __popc(inmask);
            // warp scan of `mine` -> write offsets; block totals through two shared atomics per warp
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t base = 0;
            const uint32_t wb = __reduce_add_sync(0xffffffffu, below);
            if (lane == 31) { base = atomicAdd(&sh.cursor, incl); atomicAdd(&sh.below, wb); }
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t at = base + incl - mine;
            __syncthreads();
            M = sh.cursor; B = sh.below;
            const bool low = B > k1, high = k2 >= B + M;   // the target rank lies below / above the bracket
            if (M > (uint32_t)kMonoCap || low || high) {
                if (attempt == 1 || M > (uint32_t)kMonoCap || (low && lo == 0) || (high && hi >= kExcl - 1)) return give_up(3);
                if (low) {
                    const int j = ilo - 2 * delta;
                    hi = lo - 1;
                    lo = j < 0 ? K(0) : samp[j];
                } else {
                    const int j = ihi + 2 * delta;
                    lo = hi + 1;
                    hi = j >= sv ? kExcl - 1 : samp[j];
                }
                __syncthreads();  // every thread has read the totals
                if (tid == 0) { sh.cursor = 0; sh.below = 0; }
                __syncthreads();
                continue;
            }
            gather_marked(inmask, at, std::true_type{}, lo, mono_bucket_shift<K>(span));
            break;
        }
        __syncthreads();
        mono_resolve<K, NT>(cand, M, k1 - B, k2 - B, v1k, v2k, sh, true, lo, hi);
    }

    // ================= statistics in the processed domain =================
    T m = T(0), m2 = T(0);
    T v1 = raw_val<T>(v1k), v2 = raw_val<T>(v2k);
    if (real_branch && p.norm_before) m = median_of_pair<T>(v1, v2, nv);
    // processed images of the two middle order statistics (monotone chain)
    T s1 = v1, s2 = v2;
    if (real_branch) {
        if (p.norm_before && m > T(0)) { s1 = s1 / m; s2 = s2 / m; }
        if (p.stretch != RFI_STRETCH_NONE) { s1 = apply_stretch<T>(s1, p.stretch); s2 = apply_stretch<T>(s2, p.stretch); }
        if (p.norm_after) {
            m2 = median_of_pair<T>(s1, s2, nv);
            if (m2 > T(0)) { s1 = s1 / m2; s2 = s2 / m2; }
        }
    }
    const PlanDev pp = [&]() { PlanDev q = p; if (!real_branch) { q.norm_before = q.norm_after = 0; q.stretch = RFI_STRETCH_NONE; } return q; }();
    // the extreme samples must stay finite through the chain (else: inf fill -> general kernel): two lanes
    // of warp 0 evaluate them in ONE call of the chain (not two calls by every warp), everybody reads the
    // verdict behind the next barrier
    const int pmode = proc_mode_of<T>(pp, m, m2);
    if (warp == 0) {
        const T pe = proc_mode<T>(raw_val<T>(lane == 0 ? tile_min : tile_max), pmode, m, m2);
        if (__any_sync(0xffffffffu, is_inf(pe) || is_nan(pe)) && lane == 0) sh.fail = 1;
    }

    T c = T(0), d = T(0), thr_lo = T(0), thr_hi = T(0);
    T raw_lo = T(0), raw_hi = Scalar<T>::inf();
    uint32_t nflag = 0;
    if (p.flag_mode == RFI_FLAGS_MAD) {
        c = median_of_pair<T>(s1, s2, nv);
        // ---- deviations of the sorted raw sample: V-shaped in the sample index
        K* dsamp = cand;  // [NT], free until the candidates are compacted
        bool below_c = false;
        if (tid < sv) {
            const T ps = proc_mode<T>(raw_val<T>(samp[tid]), pmode, m, m2);
            below_c = ps < c;
            dsamp[tid] = to_key<T>(fabs_(ps - c));
        }
        if (tid == 0) { sh.win[0] = sh.win[1] = sh.win[2] = sh.win[3] = -1; sh.acc[3] = 0; sh.cursor = 0; sh.below = 0; }
        __syncthreads();
        if (sh.fail) return give_up(5);
        {
            const uint32_t nb = __popc(__ballot_sync(0xffffffffu, below_c));
            if (lane == 0 && nb) atomicAdd(&sh.acc[3], nb);
        }
        __syncthreads();
        const int ju = (int)sh.acc[3];  // first sample on the upper arm (proc >= c)
        const int rho = (int)(((float)(2 * k1 + 1) * (float)sv) / (float)(2 * nv));
        const int r_in = rho - delta, r_out = rho + delta + 2;  // samples inside the inner / outer window
        if (r_in < 1 || r_out > sv - 1) return give_up(6);
        // window of r consecutive samples holding the r smallest deviations: smallest start i
        // with (i + r past the end) or (d[i] <= d[i + r] and i + r on the upper arm)
        for (int which = 0; which < 2; ++which) {
            const int r = which == 0 ? r_in : r_out;
            const int i = tid;
            if (i + r <= sv) {
                auto pred = [&](int s) {
                    if (s + r >= sv) return true;
                    return dsamp[s] <= dsamp[s + r] && (s + r) >= ju;
                };
                if (pred(i) && (i == 0 || !pred(i - 1))) { sh.win[which * 2] = i; sh.win[which * 2 + 1] = i + r - 1; }
            }
        }
        __syncthreads();
        int il = sh.win[0], iu = sh.win[1], il2 = sh.win[2], iu2 = sh.win[3];
        if (il < 0 || il2 < 0) return give_up(7);
        il2 = il2 < il ? il2 : il;
        iu2 = iu2 > iu ? iu2 : iu;
        // both windows must straddle the centre (V-shape argument)
        if (!(il2 <= il && il < ju && ju <= iu && iu <= iu2 && il2 < ju)) return give_up(8);
        // candidates reach one sample BEYOND the outer window on each side (or to the end of the
        // value range), so that everything outside is proven to deviate at least d_out
        const K L1 = samp[il], U1 = samp[iu];
        const K L2 = il2 > 0 ? samp[il2 - 1] : K(0);
        const K U2 = iu2 + 1 < sv ? samp[iu2 + 1] : kExcl - 1;
        const K d_in = dsamp[il] > dsamp[iu] ? dsamp[il] : dsamp[iu];       // interior deviations <= this
        const K d_lo2 = il2 > 0 ? dsamp[il2 - 1] : kExcl, d_up2 = iu2 + 1 < sv ? dsamp[iu2 + 1] : kExcl;
        const K d_out = d_lo2 < d_up2 ? d_lo2 : d_up2;                      // exterior deviations >= this
        // every candidate deviates at least / at most this much (each arm of the V is monotone): the range of
        // the histogram that is filled while the deviations are formed
        const K d_min = dsamp[il] < dsamp[iu] ? dsamp[il] : dsamp[iu];
        const K d_max = d_lo2 > d_up2 ? d_lo2 : d_up2;
        __syncthreads();  // dsamp (= cand) is overwritten below
        for (int b = tid; b < kMonoBuckets; b += NT) sh.hist[b] = 0;
        const K span_all = U2 - L2;
        const K w_in = U1 > L1 ? U1 - L1 - 1 : K(0);
        uint32_t allmask = 0, intmask = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            K k4[4];
            if (sizeof(K) == 4) {
                const uint4 q = *reinterpret_cast<const uint4*>(kp(g));
                k4[0] = q.x; k4[1] = q.y; k4[2] = q.z; k4[3] = q.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) k4[i] = kp(g)[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                allmask += ((K)(k4[i] - L2) <= span_all) ? (1u << (g * 4 + i)) : 0u;
                intmask += ((K)(k4[i] - L1 - 1) < w_in) ? (1u << (g * 4 + i)) : 0u;
            }
        }
        uint32_t inmask = allmask & ~intmask;   // inside the outer window, not strictly inside the inner one
        const uint32_t inside = __popc(intmask), mine = __popc(inmask);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t base = 0;
        const uint32_t wb = __reduce_add_sync(0xffffffffu, inside);
        if (lane == 31) { base = atomicAdd(&sh.cursor, incl); atomicAdd(&sh.below, wb); }
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t at = base + incl - mine;
        __syncthreads();
        const uint32_t M = sh.cursor, B = sh.below;
        if (M > (uint32_t)kMonoCap || B > k1 || k2 >= B + M) return give_up(9);
        gather_marked(inmask, at, std::false_type{}, K(0), 0);
        __syncthreads();
        // exact deviation of every candidate, in place
        const K d_width = d_max - d_min;
        const int d_shf = mono_bucket_shift<K>(d_width);
        bool outside = false;
        for (uint32_t i = tid; i < M; i += NT) {
            const T ps = proc_mode<T>(raw_val<T>(cand[i]), pmode, m, m2);
            const K dk = to_key<T>(fabs_(ps - c));
            cand[i] = dk;
            if ((K)(dk - d_min) <= d_width) atomicAdd(&sh.hist[(uint32_t)((K)(dk - d_min) >> d_shf)], 1u);
            else outside = true;
        }
        const bool hist_ok = __syncthreads_or(outside) == 0;   // (never seen to fail; the list is then scanned as before)
        K r1k, r2k;
        mono_resolve<K, NT>(cand, M, k1 - B, k2 - B, r1k, r2k, sh, hist_ok, d_min, d_max);
        // the answers must lie inside what the windows prove
        if (r1k < d_in || r2k > d_out) return give_up(11);
        d = median_of_pair<T>(from_key<T>(r1k), from_key<T>(r2k), nv);
        const T ds = d * (T)p.sigma;
        thr_hi = c + ds;
        thr_lo = c - ds;

        // ---- exact raw-domain thresholds.  Warp 0: raw_hi = max{a : proc(a) <= thr_hi}; warp 1:
        //      raw_lo = min{a : proc(a) >= thr_lo}.  The chain is inverted approximately, the 32
        //      keys around the guess are evaluated with the chain itself (one ballot); if the
        //      transition is not among them, a 32-ary search over the whole key range finds it.
        if (warp < 2) {
            constexpr K kTop = kInfKey;  // first key that is not a finite value
            const bool want_hi = (warp == 0);
            const T thr = want_hi ? thr_hi : thr_lo;
            // predicate that is true on a PREFIX of the keys: hi: proc <= thr ; lo: proc < thr
            auto pre = [&](K k) {
                const T v = proc_mode<T>(raw_val<T>(k), pmode, m, m2);
                return want_hi ? (v <= thr) : !(v >= thr);
            };
            K first_false;  // smallest key where the predicate fails (kTop if none below kTop)
            if (!(thr == thr)) first_false = want_hi ? kTop : K(0);  // NaN: nothing above / below
            else if (!pre(K(0))) first_false = 0;
            else if (pre(kTop - 1)) first_false = kTop;
            else {
                K a = 0, b = kTop - 1;  // pre(a), !pre(b)
                // approximate inverse of the chain
                T g = thr;
                if (pp.norm_after && m2 > T(0)) g = g * m2;
                if (pp.stretch == RFI_STRETCH_SQRT) g = g * g;
                else if (pp.stretch == RFI_STRETCH_LOG10) g = (T)exp10((double)g);
                if (pp.norm_before && m > T(0)) g = g * m;
                K gk = (g == g && g > T(0)) ? Scalar<T>::bits(g) : K(16);
                gk = gk < K(16) ? K(16) : (gk > kTop - 17 ? kTop - 17 : gk);
                {
                    const K t = gk - 16 + (K)lane;
                    const uint32_t bal = __ballot_sync(0xffffffffu, pre(t));
                    if ((bal & 1u) && !(bal >> 31)) {  // transition inside the window
                        const int n = __popc(bal);
                        a = gk - 16 + (K)(n - 1);
                        b = a + 1;
                    }
                }
                while (b - a > 1) {
                    const K step = (b - a + 32) / 33;
                    const K t = a + (K)(lane + 1) * step;
                    const bool good = (t < b) && pre(t);
                    const int n = __popc(__ballot_sync(0xffffffffu, good));  // monotone: a prefix of the lanes
                    const K na = a + (K)n * step, nb = a + (K)(n + 1) * step;
                    a = na;
                    b = nb < b ? nb : b;
                }
                first_false = b;
            }
            if (lane == 0) {
                if (want_hi) sh.res2 = (first_false == 0) ? kExcl : first_false - 1;  // kExcl: everything is above
                else sh.res1 = first_false;
            }
        }
        __syncthreads();
        const K klo = sh.res1, khi = sh.res2;
        // value form for phase 2: flagged iff raw < raw_lo or raw > raw_hi
        raw_lo = raw_val<T>(klo);                                   // kTop -> +inf: every finite sample
        raw_hi = (khi == kExcl) ? T(-1) : raw_val<T>(khi);          // kTop -> +inf: nothing
        // ---- flagged samples, on the raw keys
        const K khi_cmp = (khi == kExcl) ? K(0) : khi;              // "everything": key > -1
        const bool all_hi = (khi == kExcl);
        uint32_t nf = 0;
        if (nv == (uint32_t)(kP * kP) && !all_hi && klo <= khi) {
            // the common tile (no excluded key, a proper interval): flagged iff outside [klo, khi], one
            // unsigned comparison per key
            const K span_ok = khi - klo;
            uint32_t outmask = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                K k4[4];
                if (sizeof(K) == 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(kp(g));
                    k4[0] = q.x; k4[1] = q.y; k4[2] = q.z; k4[3] = q.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) k4[i] = kp(g)[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) outmask += ((K)(k4[i] - klo) > span_ok) ? (1u << (g * 4 + i)) : 0u;
            }
            nf = __popc(outmask);
        } else {
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const K x = kp(g)[i];
                    const bool f = (x != kExcl) && (all_hi || x > khi_cmp || x < klo);
                    nf += f ? 1u : 0u;
                }
            }
        }
        nf = __reduce_add_sync(0xffffffffu, nf);
        if (tid == 0) sh.acc[3] = 0;
        __syncthreads();
        if (lane == 0 && nf) atomicAdd(&sh.acc[3], nf);
        __syncthreads();
        nflag = sh.acc[3];
    } else if (p.flag_mode == RFI_FLAGS_CUSTOM) {
        __syncthreads();
        if (sh.fail) return give_up(5);
        uint32_t nf = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
            const uint32_t f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + idx));
            const uint32_t nz = (((f4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | f4) & 0x80808080u;
            nf += __popc(nz);
        }
        nf = __reduce_add_sync(0xffffffffu, nf);
        if (tid == 0) sh.acc[3] = 0;
        __syncthreads();
        if (lane == 0 && nf) atomicAdd(&sh.acc[3], nf);
        __syncthreads();
        nflag = sh.acc[3];
    } else {
        __syncthreads();
        if (sh.fail) return give_up(5);
    }

    if (tid == 0) {
        rfi_tile_stat_t st;
        st.median_before = (real_branch && p.norm_before) ? (double)m : 0.0;
        st.inf_fill = 0.0;
        st.median_after = (real_branch && p.norm_after) ? (double)m2 : 0.0;
        st.centre = (double)c; st.mad = (double)d;
        st.thr_lo = (double)thr_lo; st.thr_hi = (double)thr_hi;
        st.n_valid = (int)nv; st.n_inf = 0; st.n_flagged = (int)nflag;
        st.route = RFI_TILE_RAW_THRESHOLDS;  // monotone tile: raw thresholds valid (0 / +inf when labels are not MAD flags)
        st.raw_lo = (double)raw_lo; st.raw_hi = (double)raw_hi;
        stats[tile] = st;
        if constexpr (FUSED) *st_sh = st;
    }
    return true;
}

template <int DT, int NT, bool GK = false, int GKB = kMonoGKB, int GS = kMonoGS>
__global__ void __launch_bounds__(NT, GK ? GKB : ((sizeof(typename In<DT>::T) == 4) ? 2 : 1))
tile_stats_mono_kernel(PlanDev p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                       rfi_tile_stat_t* __restrict__ stats,
                       typename Scalar<typename In<DT>::T>::key_t* __restrict__ gkeys = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ MonoShared<typename Scalar<typename In<DT>::T>::key_t> sh;
    mono_tile_stats<DT, NT, GK, GS, false>(p, data, flags, stats, gkeys, smem_raw, sh, nullptr);
}

}  // namespace rfi
