// rfi_tiles.cu -- the create_dataset hot path on sm_100a, P = 128.
//
//   phase 1  tile_stats_mono_kernel (rfi_stats_mono.cuh)  one CTA per ORIGINAL P x P tile: load
//            (complex -> NumPy-exact magnitude fused into the 128-bit loads), exact median / MAD by
//            sampled brackets on the raw keys, flag thresholds mapped back exactly to the raw
//            domain, flagged-sample count; tile_stats_general (below, same CTA) for tiles with
//            negative / infinite / inf-filled samples or a missed bracket.  For complex input
//            through the real branch the exact magnitudes stay in the workspace for phase 2.
//   (host)   keep mask -> np.random.permutation -> dest_slot[]        (rfi_host.cpp)
//   phase 2  write_patches_kernel one CTA per original tile: log-amplitude tile in shared memory
//            (labels = two compares with the raw thresholds), gradient min/max, then every kept
//            rotation is written as (P, P, 3) float32 + (P, P) uint8 with full-sector 128-bit
//            stores, each patch once, at its final shuffled slot.
//
// Reference semantics: rfi_toolbox/preprocessing/preprocessor.py:22-42, 413-446, 562-783
// (restated in SURVEY.md Appendix A).  Statistics are rotation invariant for dims divisible
// by P, so they are computed once per original tile and shared by the R rotated patches.
#include <stdlib.h>

#include "rfi_tiles.cuh"
#include "rfi_stats_mono.cuh"

namespace rfi {

// ------------------------------------------------------------------------------------------
// phase 1
//
// The tile lives in ONE register array that holds order-preserving keys; values are
// recovered with from_key() when arithmetic is needed (the map is a bijection on non-NaN
// floats), so a select never doubles the register footprint.  Shared memory is only a
// stash for the processed values while the |x - median| keys occupy the registers.
//
// Order-statistic shortcut (identity (ii)): for non-negative finite
// input every stage (x / m with m > 0, sqrt, log10) is monotone non-decreasing after
// rounding, so the two middle order statistics of the processed tile are the images of the
// two middle order statistics of the raw tile.  One select on the raw magnitudes then
// serves the normalisation median AND the flag centre; only the MAD needs a second select.
// The shortcut is dropped (general selects) as soon as a tile has a negative, infinite or
// inf-filled sample.
// smallest key in [0, kTop] at which a prefix-true predicate fails (kTop if none), found by the calling WARP:
// 32-ary search, one ballot per step.
template <typename K, typename Pre>
RFI_DEVINL K first_false_key(Pre pre, K kTop, int lane) {
    if (!pre(K(0))) return K(0);
    if (pre(kTop - 1)) return kTop;
    K a = 0, b = kTop - 1;  // pre(a), !pre(b)
    while (b - a > 1) {
        const K step = (b - a + 32) / 33;
        const K t = a + (K)(lane + 1) * step;
        const bool good = (t < b) && pre(t);
        const int n = __popc(__ballot_sync(0xffffffffu, good));  // monotone: a prefix of the lanes
        const K na = a + (K)n * step, nb = a + (K)(n + 1) * step;
        a = na;
        b = nb < b ? nb : b;
    }
    return b;
}

template <int DT, int NT>
__device__ __noinline__ void tile_stats_general(const PlanDev& p, const void* __restrict__ data,
                                                const uint8_t* __restrict__ flags,
                                                rfi_tile_stat_t* __restrict__ stats, int route_bits,
                                                typename Scalar<typename In<DT>::T>::key_t* stash_override) {
    using T = typename In<DT>::T;
    using K = typename Scalar<T>::key_t;
    constexpr int E = kP * kP / NT;  // samples per thread
    constexpr int G = E / 4;         // groups of 4 consecutive samples
    constexpr int RS = NT / 32;      // rows covered per group step
    constexpr K kExcl = ~K(0);
    constexpr K kPosInf = to_key_const_inf<T>(false), kNegInf = to_key_const_inf<T>(true);
    constexpr K kNegZero = (K(1) << (Scalar<T>::kBits - 1)) - 1;  // key(-0.0); key(+0.0) is one above
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* stash = stash_override ? stash_override : reinterpret_cast<K*>(smem_raw);  // [E][NT], thread-private
    __shared__ BlockScratch<NT> scr;
    __shared__ RoundCounter rc;
    __shared__ SelectScratch<K> sel;
    int parity = 0, round = 0;
    round_init(rc);

    const long long tile = blockIdx.x;
    const int per = p.nh * p.nw;
    const long long w = tile / per;
    const int ti = (int)((tile % per) / p.nw), tj = (int)(tile % p.nw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t origin = ((size_t)w * p.channels + (size_t)ti * kP) * p.times + (size_t)tj * kP;

    K a[E];  // keys of the current values
#pragma unroll
    for (int g = 0; g < G; ++g) {
        size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
        T q[4];
        load4_mag<DT>(data, idx, q);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[g * 4 + i] = to_key<T>(q[i]);
    }

    // results go straight to shared memory (thread 0) so they do not occupy registers across
    // the selects; copied out once at the end
    __shared__ rfi_tile_stat_t st;
    if (threadIdx.x == 0) {
        st.median_before = st.inf_fill = st.median_after = 0.0;
        st.centre = st.mad = st.thr_lo = st.thr_hi = 0.0;
        st.n_valid = 0; st.n_inf = 0; st.n_flagged = 0; st.route = route_bits; st.raw_lo = st.raw_hi = 0.0;
    }

    // non-NaN samples (recounted before every general median: inf / inf can create a NaN)
    auto count_valid = [&]() {
        uint32_t c = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) c += (a[e] != kExcl) ? 1u : 0u;
        return round_sum(c, rc, round);
    };
    uint32_t nv, nodd;  // valid samples; samples that break the shortcut (negative or +inf)
    {
        uint32_t c = 0, o = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            c += (a[e] != kExcl) ? 1u : 0u;
            o += (a[e] < kNegZero || a[e] == kPosInf) ? 1u : 0u;
        }
        nv = round_sum(c, rc, round);
        nodd = round_sum(o, rc, round);
    }
    if (threadIdx.x == 0) st.n_valid = (int)nv;

    const bool real_branch = !In<DT>::cplx || p.magnitude;
    bool shortcut = false;  // (v1, v2) are the two middle order statistics of the current tile
    T v1 = T(0), v2 = T(0);
    // LOG10 tile whose only irregularity is that its smallest samples stretch to -inf (exact-zero bandpass rows):
    // raw thresholds are found below so that phase 2 can take its fast route (RFI_TILE_RAW_FILL)
    [[maybe_unused]] bool zero_fill = false;
    [[maybe_unused]] T zf_m = T(0);
    [[maybe_unused]] bool zf_divide = false;

    if (real_branch) {
        // ---- normalise by the median (preprocessor.py:646-670), then stretch; +-inf := MAD of
        //      the finite values (preprocessor.py:672-706).  One decode / encode per sample.
        T m = T(0);
        if (p.norm_before) {
            K k1, k2;
            m = block_median<T, NT, E>(a, nv, scr, parity, rc, round, sel, &k1, &k2);
            if (threadIdx.x == 0) st.median_before = (double)m;
            shortcut = (nodd == 0) && (nv > 0);
            v1 = from_key<T>(k1); v2 = from_key<T>(k2);
        }
        const bool divide = p.norm_before && (m > T(0));
        if (divide || p.stretch != RFI_STRETCH_NONE) {
            uint32_t ninf = 0, nfin = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                T x = from_key<T>(a[e]);
                if (divide) x = x / m;
                if (p.stretch != RFI_STRETCH_NONE) x = apply_stretch<T>(x, p.stretch);
                a[e] = to_key<T>(x);
                const bool inf = (a[e] == kPosInf || a[e] == kNegInf);
                ninf += inf ? 1u : 0u;
                nfin += (!inf && a[e] != kExcl) ? 1u : 0u;
            }
            if (divide) { v1 = v1 / m; v2 = v2 / m; }
            if (p.stretch != RFI_STRETCH_NONE) {
                v1 = apply_stretch<T>(v1, p.stretch); v2 = apply_stretch<T>(v2, p.stretch);
                ninf = round_sum(ninf, rc, round);
                if (threadIdx.x == 0) st.n_inf = (int)ninf;
                if (ninf > 0) {
                    shortcut = false;
                    nfin = round_sum(nfin, rc, round);
                    T fill = T(0);
                    if (nfin > 0) {
#pragma unroll
                        for (int e = 0; e < E; ++e) stash[e * NT + threadIdx.x] = a[e];
#pragma unroll
                        for (int e = 0; e < E; ++e) a[e] = (a[e] == kPosInf || a[e] == kNegInf) ? kExcl : a[e];
                        T c = block_median<T, NT, E>(a, nfin, scr, parity, rc, round, sel);
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            a[e] = (a[e] == kExcl) ? kExcl : to_key<T>(fabs_(from_key<T>(a[e]) - c));
                        fill = block_median<T, NT, E>(a, nfin, scr, parity, rc, round, sel);
#pragma unroll
                        for (int e = 0; e < E; ++e) a[e] = stash[e * NT + threadIdx.x];
                    }
                    if (threadIdx.x == 0) st.inf_fill = (double)fill;
                    zero_fill = sizeof(T) == 4 && p.stretch == RFI_STRETCH_LOG10 && !p.norm_after && nodd == 0 && nfin > 0;
                    zf_m = m; zf_divide = divide;
                    const K kfill = to_key<T>(fill);
#pragma unroll
                    for (int e = 0; e < E; ++e) a[e] = (a[e] == kPosInf || a[e] == kNegInf) ? kfill : a[e];
                }
            }
        }
        // ---- normalise again (preprocessor.py:309-311)
        if (p.norm_after) {
            T m2;
            if (shortcut) {
                m2 = median_of_pair<T>(v1, v2, nv);
            } else {
                const uint32_t n2 = count_valid();
                m2 = block_median<T, NT, E>(a, n2, scr, parity, rc, round, sel);
            }
            if (threadIdx.x == 0) st.median_after = (double)m2;
            if (m2 > T(0)) {
#pragma unroll
                for (int e = 0; e < E; ++e) a[e] = to_key<T>(from_key<T>(a[e]) / m2);
                v1 = v1 / m2; v2 = v2 / m2;
            }
        }
    }

    if (p.flag_mode == RFI_FLAGS_MAD) {
        // ---- MAD flags on the processed tile (preprocessor.py:708-745; |z| first for
        //      complex input, :126-127)
        uint32_t nn = nv;
        T c;
        if (shortcut) {
            c = median_of_pair<T>(v1, v2, nv);
        } else {
            nn = count_valid();
            c = block_median<T, NT, E>(a, nn, scr, parity, rc, round, sel);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            stash[e * NT + threadIdx.x] = a[e];
            a[e] = to_key<T>(fabs_(from_key<T>(a[e]) - c));
        }
        T d = block_median<T, NT, E>(a, nn, scr, parity, rc, round, sel);
        T ds = d * (T)p.sigma;
        T hi = c + ds, lo = c - ds;
        uint32_t nf = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const T x = from_key<T>(stash[e * NT + threadIdx.x]);
            nf += ((x > hi) || (x < lo)) ? 1u : 0u;
        }
        nf = round_sum(nf, rc, round);
        if (threadIdx.x == 0) {
            st.centre = (double)c; st.mad = (double)d;
            st.thr_lo = (double)lo; st.thr_hi = (double)hi;
            st.n_flagged = (int)nf;
        }
        if constexpr (sizeof(T) == 4) {
            if (zero_fill) {   // uniform
                // The samples that stretch to -inf are a PREFIX of the raw order (a / m == 0) and everything above
                // is monotone: exact raw thresholds by 32-ary searches that evaluate the chain itself, as in the
                // monotone kernel.  Phase 2 then labels a pixel by
                //     a <= raw_zero ? (fill outside [thr_lo, thr_hi]) : (a > raw_hi || a < raw_lo)
                // and takes its float32 log-amplitude chain; raw_zero travels in median_after (unused: no second
                // normalisation on such a tile).
                using RK = uint32_t;
                constexpr RK kTop = 0x7f800000u;   // first raw bit pattern that is not a finite value
                const int pmode = (zf_divide ? kProcDivM : 0) | kProcLog10;
                auto proc = [&](RK b) { return proc_mode<float>(__uint_as_float(b), pmode, (float)zf_m, 0.f); };
                __shared__ RK zsh[3];
                if (warp == 0) {
                    const RK ff = first_false_key<RK>([&](RK b) { return proc(b) == -Scalar<float>::inf(); }, kTop, lane);
                    if (lane == 0) zsh[0] = ff;
                }
                __syncthreads();
                const RK zff = zsh[0];   // first raw key whose processed value is finite
                if (zff > 0 && zff < kTop) {
                    if (warp < 2) {
                        const bool want_hi = warp == 0;
                        const float thr = want_hi ? (float)hi : (float)lo;
                        RK ff;
                        if (!(thr == thr)) ff = want_hi ? kTop : zff;   // NaN: nothing above / below
                        else ff = first_false_key<RK>([&](RK b) { if (b < zff) return true; const float v = proc(b);
                                                                  return want_hi ? (v <= thr) : !(v >= thr); }, kTop, lane);
                        if (lane == 0) zsh[want_hi ? 2 : 1] = ff;
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const RK klo = zsh[1], khi_ff = zsh[2];   // first key with proc >= thr_lo; first with proc > thr_hi
                        st.route |= RFI_TILE_RAW_THRESHOLDS | RFI_TILE_RAW_FILL;
                        st.raw_lo = (double)__uint_as_float(klo);                                   // kTop -> +inf: every finite-part sample
                        st.raw_hi = (double)__uint_as_float((khi_ff <= zff ? zff : khi_ff) - 1);    // kTop - 1: nothing
                        st.median_after = (double)__uint_as_float(zff - 1);                          // raw_zero
                    }
                    // complex input: the scratch (this algorithm's stash until here) gets the tile's exact magnitudes
                    // back, row-major as the monotone kernel leaves them, so that phase 2 reads 4 B / px there
                    if constexpr (In<DT>::cplx) {
                        if (stash_override != nullptr) {
#pragma unroll 2
                            for (int g = 0; g < G; ++g) {
                                const size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
                                T q[4];
                                load4_mag_fast<DT>(data, idx, q);
                                *reinterpret_cast<uint4*>(stash_override + ((size_t)g * NT + threadIdx.x) * 4) =
                                    make_uint4(__float_as_uint(q[0]), __float_as_uint(q[1]), __float_as_uint(q[2]), __float_as_uint(q[3]));
                            }
                        }
                    }
                }
            }
        }
    } else if (p.flag_mode == RFI_FLAGS_CUSTOM) {
        uint32_t nf = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
            uint32_t f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + idx));
            uint32_t nz = (((f4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | f4) & 0x80808080u;
            nf += __popc(nz);
        }
        nf = round_sum(nf, rc, round);
        if (threadIdx.x == 0) st.n_flagged = (int)nf;
    }
    if (threadIdx.x == 0) stats[tile] = st;
}

// the general algorithm as a kernel of its own (tests; tiles of the float64 paths reach it
// through tile_stats_mono_kernel like every other tile)
template <int DT, int NT>
__global__ void __launch_bounds__(NT, (sizeof(typename In<DT>::T) == 4) ? 2 : 1)
tile_stats_kernel(PlanDev p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                  rfi_tile_stat_t* __restrict__ stats) {
    tile_stats_general<DT, NT>(p, data, flags, stats, RFI_TILE_GENERAL, nullptr);
}

// custom flags / inference with nothing to measure on the data: flags only.
template <int NT>
__global__ void __launch_bounds__(NT)
flags_count_kernel(PlanDev p, const uint8_t* __restrict__ flags, rfi_tile_stat_t* __restrict__ stats) {
    constexpr int G = kP * kP / NT / 4, RS = NT / 32;
    __shared__ RoundCounter rc;
    int round = 0;
    round_init(rc);
    const long long tile = blockIdx.x;
    const int per = p.nh * p.nw;
    const long long w = tile / per;
    const int ti = (int)((tile % per) / p.nw), tj = (int)(tile % p.nw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t origin = ((size_t)w * p.channels + (size_t)ti * kP) * p.times + (size_t)tj * kP;
    uint32_t nf = 0;
    if (flags != nullptr) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            size_t idx = origin + (size_t)(g * RS + warp) * p.times + lane * 4;
            uint32_t f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + idx));
            uint32_t nz = (((f4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | f4) & 0x80808080u;
            nf += __popc(nz);
        }
    }
    nf = round_sum(nf, rc, round);
    if (threadIdx.x == 0) {
        rfi_tile_stat_t st;
        st.median_before = st.inf_fill = st.median_after = 0.0;
        st.centre = st.mad = st.thr_lo = st.thr_hi = 0.0;
        st.n_valid = kP * kP; st.n_inf = 0; st.n_flagged = (int)nf; st.route = 0; st.raw_lo = st.raw_hi = 0.0;
        stats[tile] = st;
    }
}

// ------------------------------------------------------------------------------------------
// phase 2
//
// Numerics of the image channels (DESIGN.md "numerics"): labels depend only on the processed
// sample and the phase-1 thresholds, and are recomputed here with the very same IEEE ops, so
// they are bit-exact.  The three image channels are tolerance-class (1e-6 relative) because
// they sit downstream of log10, which NumPy itself does not compute reproducibly; they use
//   min-max:   (v - lo) * RN(1 / range)            instead of (v - lo) / range
//   ImageNet:  fma(u, RN(1 / std), RN(-mean / std)) instead of (u - mean) / std
//   gradient:  sqrt.approx (<= 1 ulp) of td^2 + fd^2
// which keeps every value within ~2 ulp of the reference chain at a quarter of the issue slots.
template <typename T> struct Phase2Smem {
    static constexpr int kPitch = kP + 1;      // conflict-free row AND column access
    static constexpr int kFlagPitch = kP + 4;  // bytes; 33 words -> column reads hit 32 banks
};

// Position of element (row, col) of the log-amplitude tile in shared memory.  Stand-alone writer:
// pitch 129.  Single-launch kernel (SWZ): pitch 128 with the column XOR-ed by the row's low five
// bits -- the layout phase 1 left its keys in, so the tile is transformed in place; rows and
// columns are both conflict-free.
template <bool SWZ>
struct TileIdx {
    static constexpr int kPitch = SWZ ? kP : kP + 1;
    RFI_DEVINL static int at(int r, int c) { return SWZ ? (r * kP + (c ^ (r & 31))) : (r * kPitch + c); }
};

// One output row (128 px x 3 channels x float32 = 1536 contiguous bytes) leaves shared memory as ONE
// bulk asynchronous copy (cp.async.bulk, the non-tensor TMA path; SASS UBLKCP) issued by one lane,
// instead of 3 LDS.128 + 3 STG.128 per lane: the staging buffer is handed to the async proxy behind
// a proxy fence, and reused once the copy has READ it (wait_group.read).
RFI_DEVINL void bulk_store(void* dst_global, const void* src_shared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                 :: "l"(dst_global), "r"((uint32_t)__cvta_generic_to_shared(src_shared)), "r"(bytes) : "memory");
}
RFI_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
RFI_DEVINL void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Phase 2 for one original tile.  FUSED (single-launch kernel, float32 real branch): shared memory
// holds [tile 64 KB | transposed labels (in the 18 KB of phase 1's candidate list + sample) | stage];
// with `keys_in_smem` the tile region holds the exact magnitudes phase 1 left there (as bit
// patterns) and pass A is an in-place transform; otherwise (tile went through the general
// algorithm) pass A reads the cube like the stand-alone writer.
template <int DT, int NT, bool kComplexBranch, bool FUSED>
RFI_DEVINL void write_tile(const PlanDev& p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                           const rfi_tile_stat_t& st, long long tile, long long slot0, long long slot1,
                           long long slot2, long long slot3, float* __restrict__ images,
                           uint8_t* __restrict__ labels, const float* __restrict__ mag_scratch,
                           unsigned char* smem_raw, bool keys_in_smem) {
    using T = typename In<DT>::T;
    using IX = TileIdx<FUSED>;
    constexpr int E = kP * kP / NT;
    constexpr int RS = NT / 32;      // rows per step (one warp per row)
    constexpr int STEPS = kP / RS;   // row steps
    constexpr int Q = kP / 32;       // columns per lane per row (4)
    constexpr int LP = IX::kPitch;
    constexpr int FP = Phase2Smem<T>::kFlagPitch;
    static_assert(E == STEPS * Q, "tile / thread mapping");
    static_assert(!FUSED || (std::is_same<T, float>::value && !kComplexBranch && NT == kMonoNT), "fused: float32 real branch");

    T* Ls = reinterpret_cast<T*>(smem_raw);                                    // [kP][LP]
    float* Ph = reinterpret_cast<float*>(Ls + (size_t)kP * LP);               // [kP][LP] (complex branch)
    // transposed label tile (rotations 2, 3); rotations 0, 1 store their label rows from pass A
    unsigned char* FbT = reinterpret_cast<unsigned char*>(Ph + (kComplexBranch ? (size_t)kP * LP : 0));
    float* stage = reinterpret_cast<float*>(FbT + (FUSED ? (size_t)(kMonoCap + kMonoNT) * 4 : (size_t)kP * FP));  // [warps][3*kP]

    const int per = p.nh * p.nw;
    const long long w = tile / per;
    const int ti = (int)((tile % per) / p.nw), tj = (int)(tile % p.nw);
    const int R = p.rotations;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t origin = ((size_t)w * p.channels + (size_t)ti * kP) * p.times + (size_t)tj * kP;
    const T med_before = (T)st.median_before, inf_fill = (T)st.inf_fill, med_after = (T)st.median_after;
    const T thr_lo = (T)st.thr_lo, thr_hi = (T)st.thr_hi;
    const bool real_branch = !kComplexBranch;

    const float mean0 = 0.485f, mean1 = 0.456f, mean2 = 0.406f;
    const float std0 = 0.229f, std1 = 0.224f, std2 = 0.225f;
    const float is0 = 1.0f / std0, is1 = 1.0f / std1;
    const float nb0 = (0.0f - mean0) / std0, nb1 = (0.0f - mean1) / std1, nb2 = (0.0f - mean2) / std2;

    // ---- pass A: processed sample -> log amplitude tile, label tiles, min/max of L.
    // One row per warp per step; the next step's samples are prefetched into registers while
    // the current ones go through magnitude / normalise / stretch / log10 (the loop is NOT
    // unrolled over steps: the body is long and must stay inside the I-cache).
    //
    // Fast route (float32, tile measured by the monotone kernel): the label
    // is two compares of the exact magnitude with the raw-domain thresholds of phase 1 -- no
    // division, no square root -- and the log amplitude, which only feeds the tolerance-class
    // image channels, comes from the float32 chain of rfi_tiles.cuh (fast_log_amp).
    T llo = Scalar<T>::nan(), lhi = Scalar<T>::nan();
    const bool fast_route = std::is_same<T, float>::value && !kComplexBranch &&
                            (st.route & RFI_TILE_RAW_THRESHOLDS) != 0;
    // kMag (fast route, complex input through the real branch): phase 1 left the tile's exact
    // magnitudes in its scratch, row-major -- read 4 B / px from there instead of 8 B / px + |z|
    auto pass_a = [&](auto fast_tag, auto mag_tag) {
        constexpr bool kFast = decltype(fast_tag)::value;
        constexpr bool kMag = decltype(mag_tag)::value;
        using Raw = typename std::conditional<kMag, RawSample<RFI_F32>, RawSample<DT>>::type;
        constexpr int RDT = kMag ? RFI_F32 : DT;
        const void* src = kMag ? static_cast<const void*>(mag_scratch + (size_t)tile * (kP * kP)) : data;
        const size_t src_origin = kMag ? 0 : origin;
        const size_t src_pitch = kMag ? (size_t)kP : (size_t)p.times;
        [[maybe_unused]] const float raw_lo = (float)st.raw_lo, raw_hi = (float)st.raw_hi;
        // RFI_TILE_RAW_FILL (LOG10 tile whose smallest samples stretch to -inf): a sample <= raw_zero IS the fill
        // value; raw_zero travels in median_after, which such a tile does not use (tile_stats_general)
        [[maybe_unused]] const bool has_fill = (st.route & RFI_TILE_RAW_FILL) != 0;
        [[maybe_unused]] const float raw_zero = has_fill ? (float)st.median_after : -1.0f;
        [[maybe_unused]] const unsigned char fill_flag = (has_fill && (((float)inf_fill > (float)thr_hi) || ((float)inf_fill < (float)thr_lo))) ? 1 : 0;
        [[maybe_unused]] const float L_fill = log10_img(fabsf((float)inf_fill) + 1e-10f);
        [[maybe_unused]] const FastChain chain = make_fast_chain(p, (float)med_before, has_fill ? 0.0f : (float)med_after);
        Raw cur[Q], nxt[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            cur[q] = load_raw<RDT>(src, src_origin + (size_t)warp * src_pitch + lane + 32 * q);
#pragma unroll 1
        for (int s = 0; s < STEPS; ++s) {
            const int row = s * RS + warp;
            if (s + 1 < STEPS) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    nxt[q] = load_raw<RDT>(src, src_origin + (size_t)(row + RS) * src_pitch + lane + 32 * q);
            }
            unsigned char fl[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                fl[q] = 0;
                if (p.flag_mode == RFI_FLAGS_CUSTOM) fl[q] = __ldg(flags + origin + (size_t)row * p.times + lane + 32 * q);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int col = lane + 32 * q;
                T a, ph;
                unsigned char f = fl[q];
                T L;
                if constexpr (kFast) {
                    raw_to_mag_fast<RDT>(cur[q], a);
                    ph = T(0);
                    if (p.flag_mode == RFI_FLAGS_MAD) f = ((a > (T)raw_hi) || (a < (T)raw_lo)) ? 1 : 0;
                    L = (T)fast_log_amp((float)a, chain);
                    if (a <= (T)raw_zero) { f = fill_flag; L = (T)L_fill; }   // raw_zero = -1: never (a >= 0 or NaN)
                } else {
                    if constexpr (!kMag) raw_to_mag<DT, kComplexBranch>(cur[q], a, ph);
                    T x = a;
                    if (real_branch) x = process_sample<T>(a, p, med_before, inf_fill, med_after);
                    if (p.flag_mode == RFI_FLAGS_MAD) f = ((x > thr_hi) || (x < thr_lo)) ? 1 : 0;
                    L = log10_img(fabs_(x) + T(1e-10));
                }
                Ls[IX::at(row, col)] = L;
                FbT[col * FP + row] = f;
                // label rows of the untransposed rotations go out now: 32 lanes x 1 byte = one full sector
                if (slot0 >= 0) labels[(size_t)slot0 * kP * kP + (size_t)row * kP + col] = f;
                if (slot1 >= 0) labels[(size_t)slot1 * kP * kP + (size_t)(kP - 1 - row) * kP + col] = f;
                if constexpr (kComplexBranch) {
                    // (phase + pi) / (2 pi), then ImageNet (rotation invariant); tolerance class like the
                    // other channels: reciprocal multiplies and one fma instead of two IEEE divisions
                    const float c2 = (float)((ph + T(3.141592653589793)) * T(1.0 / 6.283185307179586));
                    Ph[IX::at(row, col)] = __fmaf_rn(c2, 1.0f / std2, nb2);
                } else {
                    llo = Scalar<T>::fmin_nan(llo, L);
                    lhi = Scalar<T>::fmax_nan(lhi, L);
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) cur[q] = nxt[q];
        }
    };
    // single-launch kernel: the tile region holds the exact magnitudes as bit patterns, in this very
    // layout -- every thread turns its own four keys of each row into L values IN PLACE (no barrier,
    // no second buffer) and emits the four label bytes of the row as one 32-bit word
    [[maybe_unused]] auto pass_a_inplace = [&]() {
        if constexpr (FUSED) {
            const float raw_lo = (float)st.raw_lo, raw_hi = (float)st.raw_hi;
            const FastChain chain = make_fast_chain(p, (float)med_before, (float)med_after);
            uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);
            const int px = warp & 3;  // position j of a group holds column 4 * lane + (j ^ px)
            const bool mad = p.flag_mode == RFI_FLAGS_MAD;
#pragma unroll 2
            for (int g = 0; g < STEPS; ++g) {
                const int row = g * RS + warp;
                uint32_t* slot = keys + ((size_t)g * NT + (threadIdx.x ^ ((row >> 2) & 7))) * 4;
                const uint4 kq = *reinterpret_cast<const uint4*>(slot);
                const uint32_t kk[4] = {kq.x, kq.y, kq.z, kq.w};
                float Lq[4];
                uint32_t fw = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = __uint_as_float(kk[j]);
                    const uint32_t f = (mad && ((a > raw_hi) || (a < raw_lo))) ? 1u : 0u;
                    const float L = fast_log_amp(a, chain);
                    Lq[j] = L;
                    const int i = j ^ px;
                    fw |= f << (8 * i);
                    FbT[(4 * lane + i) * FP + row] = (unsigned char)f;
                    llo = fminf(llo, L);
                    lhi = fmaxf(lhi, L);
                }
                *reinterpret_cast<float4*>(slot) = make_float4(Lq[0], Lq[1], Lq[2], Lq[3]);
                if (slot0 >= 0) reinterpret_cast<uint32_t*>(labels + (size_t)slot0 * kP * kP + (size_t)row * kP)[lane] = fw;
                if (slot1 >= 0) reinterpret_cast<uint32_t*>(labels + (size_t)slot1 * kP * kP + (size_t)(kP - 1 - row) * kP)[lane] = fw;
            }
        }
    };
    bool done = false;
    if constexpr (FUSED) {
        if (keys_in_smem) { pass_a_inplace(); done = true; }
    }
    if constexpr (DT == RFI_C64 && !kComplexBranch && !FUSED) {
        // (a tile measured by the general algorithm had its scratch used as that algorithm's stash; it is only
        //  rewritten with the magnitudes for RFI_TILE_RAW_FILL tiles)
        if (fast_route && mag_scratch != nullptr && (!(st.route & RFI_TILE_GENERAL) || (st.route & RFI_TILE_RAW_FILL))) {
            pass_a(std::true_type{}, std::true_type{}); done = true;
        }
    }
    if (!done) {
        if (fast_route) pass_a(std::true_type{}, std::false_type{});
        else pass_a(std::false_type{}, std::false_type{});
    }
    __syncthreads();

    // ---- pass B: min/max of the squared gradient for each distinct rotation variant
    T s0lo = Scalar<T>::nan(), s0hi = Scalar<T>::nan();
    T s1lo = Scalar<T>::nan(), s1hi = Scalar<T>::nan();
    T s3lo = Scalar<T>::nan(), s3hi = Scalar<T>::nan();
#pragma unroll 1
    for (int s = 0; s < STEPS; ++s) {
        const int i = s * RS + warp;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = lane + 32 * q;
            const T c = Ls[IX::at(i, j)];
            const T bi = (i > 0) ? c - Ls[IX::at(i - 1, j)] : T(0);
            const T bj = (j > 0) ? c - Ls[IX::at(i, j - 1)] : T(0);
            const T bi2 = bi * bi, bj2 = bj * bj;
            const T ss0 = Scalar<T>::fma(bi, bi, bj2);
            s0lo = Scalar<T>::fmin_nan(s0lo, ss0); s0hi = Scalar<T>::fmax_nan(s0hi, ss0);
            if (R >= 2) {
                const T fi = (i < kP - 1) ? c - Ls[IX::at(i + 1, j)] : T(0);
                const T ss1 = Scalar<T>::fma(fi, fi, bj2);
                s1lo = Scalar<T>::fmin_nan(s1lo, ss1); s1hi = Scalar<T>::fmax_nan(s1hi, ss1);
            }
            if (R >= 4) {
                const T fj = (j < kP - 1) ? c - Ls[IX::at(i, j + 1)] : T(0);
                const T ss3 = Scalar<T>::fma(fj, fj, bi2);
                s3lo = Scalar<T>::fmin_nan(s3lo, ss3); s3hi = Scalar<T>::fmax_nan(s3hi, ss3);
            }
        }
    }
    {   // all four NaN-ignoring min / max pairs in one block reduction
        T lo4[4] = {s0lo, s1lo, s3lo, llo}, hi4[4] = {s0hi, s1hi, s3hi, lhi};
        block_nanminmax4<NT, T>(lo4, hi4, reinterpret_cast<T*>(stage));
        s0lo = lo4[0]; s1lo = lo4[1]; s3lo = lo4[2]; llo = lo4[3];
        s0hi = hi4[0]; s1hi = hi4[1]; s3hi = hi4[2]; lhi = hi4[3];
    }

    // sqrt is monotone: min/max of g = sqrt(min/max of g^2)
    const ChanScale<T> g0 = make_scale<T>(sqrt_fast(s0lo), sqrt_fast(s0hi));
    const ChanScale<T> g1 = make_scale<T>(sqrt_fast(s1lo), sqrt_fast(s1hi));
    const ChanScale<T> g3 = make_scale<T>(sqrt_fast(s3lo), sqrt_fast(s3hi));
    const ChanScale<T> ls = make_scale<T>(llo, lhi);

    float* wstage = stage + (size_t)warp * 3 * kP;

    // ---- pass C: every kept rotation.  A warp owns STEPS CONSECUTIVE output rows, so the
    // row-derivative neighbour of row r is the centre value of row r-1, carried in registers;
    // only the centre and the column-derivative neighbour are read from shared memory.
    // `rot` is a compile-time constant so the source / neighbour index arithmetic folds away.
    // kPlain: both min-max channels of this rotation have a proper range (every patch but a flat or all-NaN one), so
    // the per-pixel selects that pin such a channel to exactly 0 are not compiled in -- the decision is uniform per
    // call; the real branch's third channel is a constant, staged once per warp instead of once per row.
    if constexpr (!kComplexBranch) {
#pragma unroll
        for (int q = 0; q < Q; ++q) wstage[(lane + 32 * q) * 3 + 2] = nb2;
        __syncwarp();
    }
    auto emit_as = [&](auto rot_tag, auto plain_tag, long long sl, const ChanScale<T>& gs) {
        constexpr int rot = decltype(rot_tag)::value;
        constexpr bool kPlain = decltype(plain_tag)::value;
        float* out_img = images + (size_t)sl * kP * kP * 3;
        [[maybe_unused]] unsigned char* out_lab = labels + (size_t)sl * kP * kP;
        // source element of output (orow, ocol): rot 0 (orow, ocol); rot 1 (127 - orow, ocol);
        // rot 2 (ocol, orow); rot 3 (ocol, 127 - orow).  dcol = 1: the column-derivative neighbour
        // (output column - 1)
        auto src_at = [&](int orow, int ocol, int dcol) {
            const int oc = ocol - dcol;
            if constexpr (rot == 0) return IX::at(orow, oc);
            else if constexpr (rot == 1) return IX::at(kP - 1 - orow, oc);
            else if constexpr (rot == 2) return IX::at(oc, orow);
            else return IX::at(oc, kP - 1 - orow);
        };
        const int row0 = warp * STEPS;
        // u = (v - lo) * inv, out = u / std - mean / std  ==  v * (inv / std) + (-lo * inv / std - mean / std)
        const T ga = gs.inv * (T)is0, gb = Scalar<T>::fma(-gs.lo * gs.inv, (T)is0, (T)nb0);
        [[maybe_unused]] const T la = ls.inv * (T)is1, lb = Scalar<T>::fma(-ls.lo * ls.inv, (T)is1, (T)nb1);
        T prev[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) prev[q] = (row0 > 0) ? Ls[src_at(row0 - 1, lane + 32 * q, 0)] : T(0);
#pragma unroll 1
        for (int s = 0; s < STEPS; ++s) {
            const int orow = row0 + s;  // output row i'
            float o[Q][3];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int ocol = lane + 32 * q;  // output column j'
                const int at = src_at(orow, ocol, 0);
                const T c = Ls[at];
                const T td = (orow > 0) ? c - prev[q] : T(0);
                const T fd = (ocol > 0) ? c - Ls[src_at(orow, ocol, 1)] : T(0);
                prev[q] = c;
                const T g = sqrt_fast(Scalar<T>::fma(td, td, fd * fd));
                // ((g - lo) * inv) * (1/std) - mean/std, folded; a flat (or all-NaN) channel is exactly 0
                // before the ImageNet step, whatever its pixels hold (preprocessor.py:157-163)
                o[q][0] = (kPlain || gs.ok) ? (float)Scalar<T>::fma(g, ga, gb) : nb0;
                if constexpr (kComplexBranch) {
                    T u = (c - T(-3.0)) * T(1.0 / 7.0);
                    u = u < T(0) ? T(0) : (u > T(1) ? T(1) : u);  // np.clip keeps NaN
                    o[q][1] = __fmaf_rn((float)u, is1, nb1);
                    o[q][2] = Ph[at];
                } else {
                    o[q][1] = (kPlain || ls.ok) ? (float)Scalar<T>::fma(c, la, lb) : nb1;
                    o[q][2] = nb2;
                }
            }
            // the previous row's bulk copy has read the staging buffer
            if (lane == 0) bulk_wait_read();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int ocol = lane + 32 * q;
                wstage[ocol * 3 + 0] = o[q][0];
                wstage[ocol * 3 + 1] = o[q][1];
                if constexpr (kComplexBranch) wstage[ocol * 3 + 2] = o[q][2];
            }
            fence_async_shared();
            __syncwarp();
            if (lane == 0) bulk_store(out_img + (size_t)orow * kP * 3, wstage, 3 * kP * sizeof(float));
            if constexpr (rot >= 2) {  // label row = row of the transposed label tile (flipped for rot 3)
                const unsigned char* lrow = FbT + ((rot == 2) ? orow : (kP - 1 - orow)) * FP;
                reinterpret_cast<uint32_t*>(out_lab + (size_t)orow * kP)[lane] =
                    reinterpret_cast<const uint32_t*>(lrow)[lane];
            }
        }
    };
    auto emit = [&](auto rot_tag, long long sl, const ChanScale<T>& gs) {
        if (sl < 0) return;  // uniform across the block
        if (gs.ok && (kComplexBranch || ls.ok)) emit_as(rot_tag, std::true_type{}, sl, gs);
        else emit_as(rot_tag, std::false_type{}, sl, gs);
    };
    emit(std::integral_constant<int, 0>{}, slot0, g0);
    emit(std::integral_constant<int, 1>{}, slot1, g1);
    emit(std::integral_constant<int, 2>{}, slot2, g0);
    emit(std::integral_constant<int, 3>{}, slot3, g3);
    if (lane == 0) bulk_wait_read();  // shared memory must outlive the last copy's read
}

// canonical patch index of each rotation of tile (ti, tj) of waterfall w
RFI_DEVINL void tile_slots(const PlanDev& p, const long long* __restrict__ dest_slot, long long tile,
                           long long& slot0, long long& slot1, long long& slot2, long long& slot3) {
    const int per = p.nh * p.nw;
    const long long w = tile / per;
    const int ti = (int)((tile % per) / p.nw), tj = (int)(tile % p.nw);
    const int R = p.rotations;
    const long long base = w * R * per;
    slot0 = dest_slot[base + (long long)ti * p.nw + tj];
    slot1 = slot2 = slot3 = -1;
    if (R >= 2) slot1 = dest_slot[base + per + (long long)(p.nh - 1 - ti) * p.nw + tj];
    if (R >= 4) {
        slot2 = dest_slot[base + 2LL * per + (long long)tj * p.nh + ti];
        slot3 = dest_slot[base + 3LL * per + (long long)(p.nw - 1 - tj) * p.nh + ti];
    }
}

template <int DT, int NT, bool kComplexBranch>
__global__ void __launch_bounds__(NT, (sizeof(typename In<DT>::T) == 4 && !kComplexBranch) ? 2 : 1)
write_patches_kernel(PlanDev p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                     const rfi_tile_stat_t* __restrict__ stats, const long long* __restrict__ dest_slot,
                     float* __restrict__ images, uint8_t* __restrict__ labels,
                     const float* __restrict__ mag_scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long tile = blockIdx.x;
    long long slot0, slot1, slot2, slot3;
    tile_slots(p, dest_slot, tile, slot0, slot1, slot2, slot3);
    if (slot0 < 0 && slot1 < 0 && slot2 < 0 && slot3 < 0) return;
    const rfi_tile_stat_t st = stats[tile];
    write_tile<DT, NT, kComplexBranch, false>(p, data, flags, st, tile, slot0, slot1, slot2, slot3, images, labels,
                                              mag_scratch, smem_raw, false);
}

// ------------------------------------------------------------------------------------------
// phase 1 + phase 2 in ONE launch (float32 arithmetic, real branch), for a caller that knows every
// patch's destination slot before the statistics exist: inference_mode (no compaction, no shuffle)
// or MAD flags with the slots of the all-kept case drawn ahead (the Python layer checks the flag
// counts afterwards and falls back to the two-phase path if a tile turned out blank).  One CTA per
// original tile, 2 CTAs / SM: while one CTA sits in the barrier-bound selection, the other streams
// its 4 x 196 KB of patches out -- the issue-bound and the HBM-bound halves of the path overlap
// inside every SM, the cube is read once (8 B / px) and nothing but the output is written: 60 B / px.
constexpr size_t kFusedTileBytes = (size_t)kP * kP * 4;
constexpr size_t kFusedAuxBytes = (size_t)(kMonoCap + kMonoNT) * 4;      // candidates + sample | transposed labels
constexpr size_t kFusedStageBytes = (size_t)(kMonoNT / 32) * 3 * kP * 4;  // staging rows | phase 1's MonoShared
constexpr size_t kFusedSmem = kFusedTileBytes + kFusedAuxBytes + kFusedStageBytes;
static_assert(sizeof(MonoShared<uint32_t>) <= kFusedStageBytes, "MonoShared aliases the staging rows");
static_assert((size_t)kP * Phase2Smem<float>::kFlagPitch <= kFusedAuxBytes, "transposed labels alias phase 1's lists");

template <int DT>
__global__ void __launch_bounds__(kMonoNT, 2)
tile_fused_kernel(PlanDev p, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                  rfi_tile_stat_t* __restrict__ stats, const long long* __restrict__ dest_slot,
                  float* __restrict__ images, uint8_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ rfi_tile_stat_t st_sh;
    MonoShared<uint32_t>& sh = *reinterpret_cast<MonoShared<uint32_t>*>(smem_raw + kFusedTileBytes + kFusedAuxBytes);
    const long long tile = blockIdx.x;
    const bool mono = mono_tile_stats<DT, kMonoNT, false, kMonoGS, true>(p, data, flags, stats, nullptr, smem_raw, sh, &st_sh);
    __syncthreads();  // statistics visible (shared copy, or stats[tile] written by this CTA's thread 0)
    long long slot0, slot1, slot2, slot3;
    tile_slots(p, dest_slot, tile, slot0, slot1, slot2, slot3);
    if (slot0 < 0 && slot1 < 0 && slot2 < 0 && slot3 < 0) return;
    const rfi_tile_stat_t st = mono ? st_sh : stats[tile];
    write_tile<DT, kMonoNT, false, true>(p, data, flags, st, tile, slot0, slot1, slot2, slot3, images, labels,
                                         nullptr, smem_raw, mono);
}

// ------------------------------------------------------------------------------------------
// host side
static int make_plan(const rfi_plan_t* plan, PlanDev& d) {
    if (!plan) { set_error("plan is NULL"); return RFI_E_INVALID; }
    if (plan->dtype < RFI_F32 || plan->dtype > RFI_C128) { set_error("bad dtype %d", plan->dtype); return RFI_E_INVALID; }
    if (plan->rotations != 1 && plan->rotations != 2 && plan->rotations != 4) {
        set_error("rotations must be 1, 2 or 4 (got %d)", plan->rotations); return RFI_E_INVALID; }
    if (plan->stretch < 0 || plan->stretch > 2) { set_error("bad stretch %d", plan->stretch); return RFI_E_INVALID; }
    if (plan->flag_mode < 0 || plan->flag_mode > 2) { set_error("bad flag_mode %d", plan->flag_mode); return RFI_E_INVALID; }
    if (plan->n_waterfalls < 0 || plan->channels <= 0 || plan->times <= 0) { set_error("bad cube shape"); return RFI_E_INVALID; }
    if (plan->patch <= 0) { set_error("bad patch size %d", plan->patch); return RFI_E_INVALID; }
    if (!plan_is_fast(plan)) { set_error("internal: plan is not a fast-path plan"); return RFI_E_INVALID; }
    d.n_waterfalls = plan->n_waterfalls; d.channels = plan->channels; d.times = plan->times;
    d.nh = (int)(plan->channels / kP); d.nw = (int)(plan->times / kP);
    d.rotations = plan->rotations; d.stretch = plan->stretch;
    d.norm_before = plan->norm_before; d.norm_after = plan->norm_after;
    d.flag_mode = plan->flag_mode; d.magnitude = plan->magnitude; d.sigma = plan->sigma;
    return RFI_OK;
}

template <int DT, int NT>
static int launch_stats(const PlanDev& d, long long tiles, const void* data, const uint8_t* flags,
                        rfi_tile_stat_t* stats, void* scratch, cudaStream_t st) {
    using K = typename Scalar<typename In<DT>::T>::key_t;
    if constexpr (sizeof(K) == 4) {
        // float32 keys: half of them in the thread-private global scratch, 3 CTAs / SM
        if (!scratch) { set_error("rfi_tile_stats needs the workspace rfi_plan_workspace_bytes() reports"); return RFI_E_INVALID; }
        // experiment knobs (co-residency with the writer): RFI_MONO_GS=3 keeps 3 of the 8 key
        // groups in shared memory (45 KB / CTA instead of 57 KB); RFI_STATS_SMEM_PAD_KB pads the request
        static const int gs_env = getenv("RFI_MONO_GS") ? atoi(getenv("RFI_MONO_GS")) : kMonoGS;
        static const int pad_kb = getenv("RFI_STATS_SMEM_PAD_KB") ? atoi(getenv("RFI_STATS_SMEM_PAD_KB")) : 0;
        if (gs_env == 3) {
            auto k = tile_stats_mono_kernel<DT, kMonoNT, true, kMonoGKB, 3>;
            const size_t gsmem = (size_t)(kMonoCap + kMonoNT + 3 * kMonoNT * 4) * sizeof(K) + (size_t)pad_kb * 1024;
            RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            k<<<(unsigned)tiles, kMonoNT, gsmem, st>>>(d, data, flags, stats, static_cast<K*>(scratch));
        } else {
            auto k = tile_stats_mono_kernel<DT, kMonoNT, true>;
            const size_t gsmem = (size_t)(kMonoCap + kMonoNT + kMonoGS * kMonoNT * 4) * sizeof(K) + (size_t)pad_kb * 1024;
            RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            k<<<(unsigned)tiles, kMonoNT, gsmem, st>>>(d, data, flags, stats, static_cast<K*>(scratch));
        }
    } else {
        auto mono = tile_stats_mono_kernel<DT, kMonoNT>;
        size_t msmem = (size_t)(kP * kP + kMonoCap + kMonoNT) * sizeof(K);
        RFI_CUDA_TRY(cudaFuncSetAttribute(mono, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
        mono<<<(unsigned)tiles, kMonoNT, msmem, st>>>(d, data, flags, stats, nullptr);
    }
    // tiles with negative / infinite / inf-filled samples, or whose sampled bracket missed, run
    // the general algorithm inside the same CTA (no second launch, no tail)
    static_assert(NT == kMonoNT, "the general path runs inside the monotone kernel's CTA");
    return RFI_OK;
}

template <int DT, int NT, bool CB>
static int launch_write(const PlanDev& d, long long tiles, const void* data, const uint8_t* flags,
                        const rfi_tile_stat_t* stats, const long long* dest, float* images,
                        uint8_t* labels, const float* mag_scratch, cudaStream_t st) {
    using T = typename In<DT>::T;
    auto kern = write_patches_kernel<DT, NT, CB>;
    static const int pad_kb = getenv("RFI_WRITER_SMEM_PAD_KB") ? atoi(getenv("RFI_WRITER_SMEM_PAD_KB")) : 0;  // experiment knob
    size_t smem = (size_t)kP * Phase2Smem<T>::kPitch * sizeof(T) +
                  (CB ? (size_t)kP * Phase2Smem<T>::kPitch * sizeof(float) : 0) +
                  (size_t)kP * Phase2Smem<T>::kFlagPitch + (size_t)(NT / 32) * (3 * kP * sizeof(float)) +
                  (size_t)pad_kb * 1024;
    RFI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)tiles, NT, smem, st>>>(d, data, flags, stats, dest, images, labels, mag_scratch);
    return RFI_OK;
}

}  // namespace rfi

using namespace rfi;

extern "C" int64_t rfi_plan_num_tiles(const rfi_plan_t* plan) {
    if (!plan || plan->patch <= 0) return -1;
    if (plan_is_fast(plan)) return plan->n_waterfalls * (plan->channels / plan->patch) * (plan->times / plan->patch);
    return generic_num_groups(plan);
}
extern "C" int64_t rfi_plan_num_patches(const rfi_plan_t* plan) {
    if (!plan || plan->patch <= 0) return -1;
    if (plan_is_fast(plan)) return rfi_plan_num_tiles(plan) * plan->rotations;
    return generic_num_patches(plan);
}
extern "C" int rfi_plan_path(const rfi_plan_t* plan) {
    if (!plan || plan->patch <= 0) return -1;
    return plan_is_fast(plan) ? RFI_PATH_FAST : plan_is_big(plan) ? RFI_PATH_BIG : RFI_PATH_GENERIC;
}
extern "C" size_t rfi_plan_workspace_bytes(const rfi_plan_t* plan) {
    if (!plan || plan->patch <= 0) return 0;
    if (plan_is_fast(plan)) {
        // float32 arithmetic: the statistics kernel keeps part of every tile's keys in a
        // thread-private scratch (64 KB per tile = 4 B / px); float64 tiles stay on chip
        if (plan->dtype != RFI_F32 && plan->dtype != RFI_C64) return 0;
        return (size_t)rfi_plan_num_tiles(plan) * kP * kP * sizeof(uint32_t);
    }
    if (plan_is_big(plan)) return big_workspace_bytes(plan);
    return generic_workspace_bytes(plan);
}

extern "C" int rfi_tile_stats(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                              rfi_tile_stat_t* stats, void* workspace, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (plan && plan->patch > 0 && !plan_is_fast(plan)) {
        if (plan_is_big(plan)) return big_tile_stats(plan, data, flags, stats, workspace, st);
        return generic_tile_stats(plan, data, flags, stats, workspace, st);
    }
    PlanDev d;
    int rc = make_plan(plan, d);
    if (rc) return rc;
    const long long tiles = rfi_plan_num_tiles(plan);
    if (tiles == 0) return RFI_OK;
    if (!data || !stats) { set_error("data / stats is NULL"); return RFI_E_INVALID; }
    if (d.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    const bool cplx = plan->dtype >= RFI_C64;
    const bool real_branch = !cplx || plan->magnitude;
    const bool need_data = d.flag_mode == RFI_FLAGS_MAD ||
                           (real_branch && (d.norm_before || d.norm_after || d.stretch != RFI_STRETCH_NONE));
    if (!need_data) {
        flags_count_kernel<512><<<(unsigned)tiles, 512, 0, st>>>(d, d.flag_mode == RFI_FLAGS_CUSTOM ? flags : nullptr, stats);
    } else {
        switch (plan->dtype) {
            case RFI_F32: rc = launch_stats<RFI_F32, 512>(d, tiles, data, flags, stats, workspace, st); break;
            case RFI_C64: rc = launch_stats<RFI_C64, 512>(d, tiles, data, flags, stats, workspace, st); break;
            case RFI_F64: rc = launch_stats<RFI_F64, 512>(d, tiles, data, flags, stats, workspace, st); break;
            default:      rc = launch_stats<RFI_C128, 512>(d, tiles, data, flags, stats, workspace, st); break;
        }
        if (rc) return rc;
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

// ---- single launch (see tile_fused_kernel) ---------------------------------------------------
static bool plan_is_fusable(const rfi_plan_t* plan) {
    if (!plan || plan->patch <= 0 || !plan_is_fast(plan)) return false;
    const bool f32_real = plan->dtype == RFI_F32 || (plan->dtype == RFI_C64 && plan->magnitude);
    return f32_real && (plan->flag_mode == RFI_FLAGS_MAD || plan->flag_mode == RFI_FLAGS_INFERENCE);
}
extern "C" int rfi_plan_fusable(const rfi_plan_t* plan) { return plan_is_fusable(plan) ? 1 : 0; }

extern "C" int rfi_fused_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                                 rfi_tile_stat_t* stats, const int64_t* dest_slot, float* images,
                                 uint8_t* labels, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!plan_is_fusable(plan)) { set_error("plan cannot take the single-launch path (rfi_plan_fusable)"); return RFI_E_UNSUPPORTED; }
    PlanDev d;
    int rc = make_plan(plan, d);
    if (rc) return rc;
    const long long tiles = rfi_plan_num_tiles(plan);
    if (tiles == 0) return RFI_OK;
    if (!data || !stats || !dest_slot) { set_error("data / stats / dest_slot is NULL"); return RFI_E_INVALID; }
    const long long* dest = reinterpret_cast<const long long*>(dest_slot);
    if (plan->dtype == RFI_F32) {
        auto k = tile_fused_kernel<RFI_F32>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
        k<<<(unsigned)tiles, kMonoNT, kFusedSmem, st>>>(d, data, flags, stats, dest, images, labels);
    } else {
        auto k = tile_fused_kernel<RFI_C64>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
        k<<<(unsigned)tiles, kMonoNT, kFusedSmem, st>>>(d, data, flags, stats, dest, images, labels);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

extern "C" int rfi_write_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                                 const rfi_tile_stat_t* stats, const int64_t* dest_slot,
                                 float* images, uint8_t* labels, void* workspace, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long* dest = reinterpret_cast<const long long*>(dest_slot);
    if (plan && plan->patch > 0 && !plan_is_fast(plan)) {
        if (plan_is_big(plan)) return big_write_patches(plan, data, flags, stats, dest, images, labels, workspace, st);
        return generic_write_patches(plan, data, flags, stats, dest, images, labels, workspace, st);
    }
    PlanDev d;
    int rc = make_plan(plan, d);
    if (rc) return rc;
    const long long tiles = rfi_plan_num_tiles(plan);
    if (tiles == 0) return RFI_OK;
    if (!data || !stats || !dest_slot) { set_error("data / stats / dest_slot is NULL"); return RFI_E_INVALID; }
    if (d.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    const bool cb = plan->dtype >= RFI_C64 && !plan->magnitude;
    // complex64 through the real branch: the statistics kernel left every monotone tile's exact
    // magnitudes in the workspace (row-major per tile); NULL = read the cube again
    const bool need_data = d.flag_mode == RFI_FLAGS_MAD || (d.norm_before || d.norm_after || d.stretch != RFI_STRETCH_NONE);
    const float* mag = (plan->dtype == RFI_C64 && plan->magnitude && need_data) ? static_cast<const float*>(workspace) : nullptr;
    switch (plan->dtype) {
        case RFI_F32: rc = launch_write<RFI_F32, 512, false>(d, tiles, data, flags, stats, dest, images, labels, mag, st); break;
        case RFI_F64: rc = launch_write<RFI_F64, 512, false>(d, tiles, data, flags, stats, dest, images, labels, mag, st); break;
        case RFI_C64:
            rc = cb ? launch_write<RFI_C64, 1024, true>(d, tiles, data, flags, stats, dest, images, labels, mag, st)
                    : launch_write<RFI_C64, 512, false>(d, tiles, data, flags, stats, dest, images, labels, mag, st);
            break;
        default:
            rc = cb ? launch_write<RFI_C128, 256, true>(d, tiles, data, flags, stats, dest, images, labels, mag, st)
                    : launch_write<RFI_C128, 512, false>(d, tiles, data, flags, stats, dest, images, labels, mag, st);
            break;
    }
    if (rc) return rc;
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
