// rfi_pairs.cu -- BASELINE config 4 on sm_100a: compute_ffi + MAD / std reduction + the IoU / F1
// confusion counts of every (data, predicted mask, true mask) patch pair in ONE launch, every byte
// of the three inputs read once (8 + 1 + 1 = 10 B / px for complex64).
//
// Replaces, per pair, rfi_toolbox/evaluation/statistics.py:10-97 (compute_statistics before and
// after flagging, compute_ffi) and the boolean reductions of evaluation/metrics.py:36-40, 63-68,
// 95-99, 142-147 behind evaluate_segmentation (:155-172).
//
// One CTA of 512 threads per pair (<= 16384 samples, e.g. one 128 x 128 patch), 2 CTAs / SM:
//   load     128-bit loads, NumPy-exact |z|, order-preserving keys to shared memory (64 KB);
//            the pair's flag bytes become one 32-bit mask per thread, the same loads feed
//            TP / FP / FN; float64 sums for the two means
//   moments  squared deviations from the T-rounded means (np.std's own formula), both sets
//   select   median and MAD of ALL samples, then of the UNFLAGGED samples (flagged keys are
//            rewritten to the excluded key): the sampled-bracket selection of phase 1
//            (rfi_stats_mono.cuh: sorted 512-sample, 3 sigma rank brackets, one counting sweep,
//            one compaction sweep, exact rank by mono_resolve; the MAD through the V-shaped
//            deviation windows of the sorted sample) with every bracket validated; a pair whose
//            bracket misses (or with < 64 samples, or +-inf samples) goes through a 4-bit radix
//            select over the same shared-memory keys -- results are exact order statistics either way
//   result   the FFI arithmetic of statistics.py:77-97 in float64, operation by operation (the
//            build forbids FMA contraction), so the values equal the reference's Python floats
#include "rfi_stats_mono.cuh"

namespace rfi {

constexpr int kPairNT = 512, kPairE = 32, kPairG = kPairE / 4, kPairSeg = kPairNT * kPairE;
using PK = uint32_t;
constexpr PK kPExcl = ~PK(0);

struct PairShared {
    MonoShared<PK> ms;
    double pd[4][16];     // per-warp partials
    uint32_t pu[8][16];
    double rd[4];         // block totals
    uint32_t ru[8];
    uint32_t cnt[4][16];  // radix-select counters, rotating (see RoundCounter)
    uint32_t nxt[2];      // count(key <= prefix), min(key > prefix); the two maxima of the load pass
    uint32_t ext[2];      // the two minima of the load pass
};

RFI_DEVINL uint4 pair_keys(const PK* skeys, int g) {
    return *reinterpret_cast<const uint4*>(skeys + ((size_t)g * kPairNT + threadIdx.x) * 4);
}

// ---- deterministic block totals: warp shuffles, one slot per warp, warp 0 folds the 16 slots
template <int ND, int NU>
RFI_DEVINL void pair_totals(double (&d)[ND], uint32_t (&u)[NU], PairShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
#pragma unroll
        for (int o = 16; o; o >>= 1) d[i] += __shfl_xor_sync(0xffffffffu, d[i], o);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = __reduce_add_sync(0xffffffffu, u[i]);
    __syncthreads();  // previous readers of the totals are done
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < ND; ++i) sh.pd[i][warp] = d[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) sh.pu[i][warp] = u[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            double v = lane < 16 ? sh.pd[i][lane] : 0.0;
#pragma unroll
            for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sh.rd[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            uint32_t v = lane < 16 ? sh.pu[i][lane] : 0u;
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) sh.ru[i] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ND; ++i) d[i] = sh.rd[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = sh.ru[i];
}

// ---- exact fallback: MSB-first radix select over the shared-memory keys, 4 bits per pass, 15
// register counters per thread (no histogram atomics).  DEV: the key of a sample is its absolute
// deviation from `centre`.  Returns the keys of ranks k1 <= k2 <= k1 + 1 among the valid keys.
template <bool DEV>
__device__ __noinline__ void pair_radix_select(const PK* __restrict__ skeys, uint32_t k1, uint32_t k2, float centre,
                                               PK& o1, PK& o2, PairShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31;
    auto key_of = [&](PK x) -> PK {
        if (!DEV) return x;
        return x == kPExcl ? kPExcl : to_key<float>(fabsf(from_key<float>(x) - centre));
    };
    if (tid < 64) sh.cnt[tid >> 4][tid & 15] = 0;
    __syncthreads();
    PK prefix = 0;
    int round = 0;
#pragma unroll 1
    for (int shift = 28; shift >= 0; shift -= 4, ++round) {
        uint32_t c[15];
#pragma unroll
        for (int t = 0; t < 15; ++t) c[t] = 0;
#pragma unroll 2
        for (int g = 0; g < kPairG; ++g) {
            const uint4 q = pair_keys(skeys, g);
            const PK k4[4] = {key_of(q.x), key_of(q.y), key_of(q.z), key_of(q.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int t = 0; t < 15; ++t) c[t] += (k4[i] < (prefix | ((PK)(t + 1) << shift))) ? 1u : 0u;
            }
        }
        uint32_t* slot = sh.cnt[round & 3];
#pragma unroll
        for (int t = 0; t < 15; ++t) {
            const uint32_t v = __reduce_add_sync(0xffffffffu, c[t]);
            if (lane == 0 && v) atomicAdd(&slot[t], v);
        }
        __syncthreads();
        int d = 0;
#pragma unroll
        for (int t = 0; t < 15; ++t) d += (slot[t] <= k1) ? 1 : 0;
        prefix |= (PK)d << shift;
        if (tid < 16) sh.cnt[(round + 2) & 3][tid] = 0;  // last read before the previous barrier
    }
    o1 = o2 = prefix;
    if (k2 != k1) {  // rank k1 + 1: the same key if duplicates reach it, else the smallest key above
        uint32_t cle = 0;
        PK nxt = kPExcl;
#pragma unroll 2
        for (int g = 0; g < kPairG; ++g) {
            const uint4 q = pair_keys(skeys, g);
            const PK k4[4] = {key_of(q.x), key_of(q.y), key_of(q.z), key_of(q.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cle += (k4[i] <= prefix) ? 1u : 0u;
                const PK y = k4[i] > prefix ? k4[i] : kPExcl;
                nxt = y < nxt ? y : nxt;
            }
        }
        cle = __reduce_add_sync(0xffffffffu, cle);
        nxt = warp_min(nxt);
        __syncthreads();
        if (tid == 0) { sh.nxt[0] = 0; sh.nxt[1] = kPExcl; }
        __syncthreads();
        if (lane == 0) { atomicAdd(&sh.nxt[0], cle); atomicMin(&sh.nxt[1], nxt); }
        __syncthreads();
        if (k2 >= sh.nxt[0]) o2 = sh.nxt[1];
    }
    __syncthreads();
}

// ---- sorted 512-sample of the valid keys (one element per thread, stratified like phase 1's):
// four warps sort 128 keys each in registers, every thread ranks one key in the other three runs.
// `runs` is scratch (the candidate list).  Returns the number of valid samples (sorted first).
RFI_DEVINL int pair_sort_sample(const PK* __restrict__ skeys, PK* runs, PK* samp) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int e_s = (((lane + warp * 3) & 7) << 2) | (((lane >> 3) + warp) & 3);
    PK x = skeys[((size_t)(e_s >> 2) * kPairNT + tid) * 4 + (e_s & 3)];
    runs[tid] = x;
    const int nvalid = __syncthreads_count(x != kPExcl);
    if (warp < 4) {
        PK v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = runs[warp * 128 + r * 32 + lane];
        warp_sort_regs<PK, 4>(v, lane);
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
        const PK* run = runs + o * 128;
        const bool incl = o < run_id;  // earlier runs win ties
        uint32_t pos = 0;
#pragma unroll
        for (int step = 64; step > 0; step >>= 1) {
            const PK y = run[pos + step - 1];
            pos += (incl ? (y <= x) : (y < x)) ? step : 0;
        }
        const PK y = run[127];
        pos += (pos == 127 && (incl ? (y <= x) : (y < x))) ? 1u : 0u;
        rank += pos;
    }
    samp[rank] = x;
    __syncthreads();
    return nvalid;
}

// ---- two middle order statistics of the nv valid keys by a sampled bracket (one directional retry).
// false = bracket missed / too many candidates: the caller runs the radix select.
__device__ __noinline__ bool pair_sampled_median(const PK* __restrict__ skeys, PK* cand, const PK* samp,
                                                 PairShared& shp, uint32_t nv, int sv, PK& v1k, PK& v2k) {
    MonoShared<PK>& sh = shp.ms;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    const int delta = (int)(kMonoSigma * 0.5f * sqrtf((float)sv)) + 2;
    const int rho = (int)(((float)(2 * k1 + 1) * (float)sv) / (float)(2 * nv));
    const int ilo = rho - delta, ihi = rho + delta + 1;
    PK lo = ilo < 0 ? PK(0) : samp[ilo];
    PK hi = ihi >= sv ? kPExcl - 1 : samp[ihi];
    if (tid == 0) { sh.cursor = 0; sh.below = 0; }
    __syncthreads();
    uint32_t M = 0, B = 0;
#pragma unroll 1
    for (int attempt = 0;; ++attempt) {
        const PK span = hi - lo;
        uint32_t below = 0, mine = 0;
#pragma unroll
        for (int g = 0; g < kPairG; ++g) {
            const uint4 q = pair_keys(skeys, g);
            const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                below += (k4[i] < lo) ? 1u : 0u;
                mine += ((PK)(k4[i] - lo) <= span) ? 1u : 0u;
            }
        }
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
            const bool dead = attempt == 1 || M > (uint32_t)kMonoCap || (low && lo == 0) || (high && hi >= kPExcl - 1);
            __syncthreads();  // every thread has read the totals
            if (dead) return false;
            if (low) {
                const int j = ilo - 2 * delta;
                hi = lo - 1;
                lo = j < 0 ? PK(0) : samp[j];
            } else {
                const int j = ihi + 2 * delta;
                lo = hi + 1;
                hi = j >= sv ? kPExcl - 1 : samp[j];
            }
            if (tid == 0) { sh.cursor = 0; sh.below = 0; }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int g = 0; g < kPairG; ++g) {
            const uint4 q = pair_keys(skeys, g);
            const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if ((PK)(k4[i] - lo) <= span) cand[at++] = k4[i];
        }
        break;
    }
    __syncthreads();
    mono_resolve<PK, kPairNT>(cand, M, k1 - B, k2 - B, v1k, v2k, sh);
    return true;
}

// ---- two middle order statistics of |x - c| over the valid keys: |x - c| is V-shaped in x, so the
// r smallest deviations are a contiguous window of the sorted sample around c; an inner and an outer
// window prove bounds for everything strictly inside / outside, only the candidates in between get
// their exact deviation, and the answer is accepted only inside what was proven.
__device__ __noinline__ bool pair_sampled_mad(const PK* __restrict__ skeys, PK* cand, const PK* samp,
                                              PairShared& shp, uint32_t nv, int sv, float c, PK& r1k, PK& r2k) {
    MonoShared<PK>& sh = shp.ms;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    const int delta = (int)(kMonoSigma * 0.5f * sqrtf((float)sv)) + 2;
    PK* dsamp = cand;  // [NT], free until the candidates are compacted
    bool below_c = false;
    if (tid < sv) {
        const float ps = from_key<float>(samp[tid]);
        below_c = ps < c;
        dsamp[tid] = to_key<float>(fabsf(ps - c));
    }
    if (tid == 0) { sh.win[0] = sh.win[1] = sh.win[2] = sh.win[3] = -1; sh.acc[3] = 0; sh.cursor = 0; sh.below = 0; }
    __syncthreads();
    {
        const uint32_t nb = __popc(__ballot_sync(0xffffffffu, below_c));
        if (lane == 0 && nb) atomicAdd(&sh.acc[3], nb);
    }
    __syncthreads();
    const int ju = (int)sh.acc[3];  // first sample on the upper arm (value >= c)
    const int rho = (int)(((float)(2 * k1 + 1) * (float)sv) / (float)(2 * nv));
    const int r_in = rho - delta, r_out = rho + delta + 2;  // samples inside the inner / outer window
    if (r_in < 1 || r_out > sv - 1) { __syncthreads(); return false; }
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
    if (il < 0 || il2 < 0) { __syncthreads(); return false; }
    il2 = il2 < il ? il2 : il;
    iu2 = iu2 > iu ? iu2 : iu;
    if (!(il2 <= il && il < ju && ju <= iu && iu <= iu2 && il2 < ju)) { __syncthreads(); return false; }
    const PK L1 = samp[il], U1 = samp[iu];
    const PK L2 = il2 > 0 ? samp[il2 - 1] : PK(0);
    const PK U2 = iu2 + 1 < sv ? samp[iu2 + 1] : kPExcl - 1;
    const PK d_in = dsamp[il] > dsamp[iu] ? dsamp[il] : dsamp[iu];       // interior deviations <= this
    const PK d_lo2 = il2 > 0 ? dsamp[il2 - 1] : kPExcl, d_up2 = iu2 + 1 < sv ? dsamp[iu2 + 1] : kPExcl;
    const PK d_out = d_lo2 < d_up2 ? d_lo2 : d_up2;                      // exterior deviations >= this
    __syncthreads();  // dsamp (= cand) is overwritten below
    const PK span_all = U2 - L2;
    const PK w_in = U1 > L1 ? U1 - L1 - 1 : PK(0);
    uint32_t inside = 0, mine = 0;
#pragma unroll
    for (int g = 0; g < kPairG; ++g) {
        const uint4 q = pair_keys(skeys, g);
        const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool in_all = (PK)(k4[i] - L2) <= span_all;
            const bool interior = (PK)(k4[i] - L1 - 1) < w_in;
            inside += interior ? 1u : 0u;
            mine += (in_all && !interior) ? 1u : 0u;
        }
    }
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
    if (M > (uint32_t)kMonoCap || B > k1 || k2 >= B + M) { __syncthreads(); return false; }
#pragma unroll
    for (int g = 0; g < kPairG; ++g) {
        const uint4 q = pair_keys(skeys, g);
        const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool in_all = (PK)(k4[i] - L2) <= span_all;
            const bool interior = (PK)(k4[i] - L1 - 1) < w_in;
            if (in_all && !interior) cand[at++] = k4[i];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < M; i += kPairNT) cand[i] = to_key<float>(fabsf(from_key<float>(cand[i]) - c));
    __syncthreads();
    mono_resolve<PK, kPairNT>(cand, M, k1 - B, k2 - B, r1k, r2k, sh);
    return !(r1k < d_in || r2k > d_out);  // the answers must lie inside what the windows prove
}

// median and MAD of the nv valid keys (nv >= 1, no NaN among them)
__device__ __noinline__ void pair_median_mad(const PK* __restrict__ skeys, PK* cand, PK* samp, PairShared& sh,
                                             uint32_t nv, bool any_inf, float& med, float& mad) {
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    bool sampled = nv >= 64 && !any_inf;
    int sv = 0;
    if (sampled) {
        sv = pair_sort_sample(skeys, cand, samp);
        sampled = sv >= 64;
    }
    PK a = 0, b = 0;
    if (!(sampled && pair_sampled_median(skeys, cand, samp, sh, nv, sv, a, b)))
        pair_radix_select<false>(skeys, k1, k2, 0.f, a, b, sh);
    med = median_of_pair<float>(from_key<float>(a), from_key<float>(b), nv);
    if (is_inf(med) || is_nan(med)) {  // some |x - med| is inf - inf = NaN: np.median propagates it
        mad = Scalar<float>::nan();
        return;
    }
    if (!(sampled && pair_sampled_mad(skeys, cand, samp, sh, nv, sv, med, a, b)))
        pair_radix_select<true>(skeys, k1, k2, med, a, b, sh);
    mad = median_of_pair<float>(from_key<float>(a), from_key<float>(b), nv);
}

template <int DT>
__global__ void __launch_bounds__(kPairNT, 2)
pair_sweep_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, const uint8_t* __restrict__ truth,
                  long long seg, rfi_stats_t* __restrict__ out, rfi_pair_result_t* __restrict__ res) {
    static_assert(DT == RFI_F32 || DT == RFI_C64, "float32 arithmetic");
    constexpr int NT = kPairNT, G = kPairG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PK* skeys = reinterpret_cast<PK*>(smem_raw);   // [G][NT][4]
    PK* cand = skeys + kPairSeg;                   // [kMonoCap]
    PK* samp = cand + kMonoCap;                    // [NT]
    __shared__ PairShared sh;
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * seg;
    const uint32_t n_all = (uint32_t)seg;
    const size_t esz = DT == RFI_C64 ? 8 : 4;
    const bool vec = ((seg & 3) == 0) && ((reinterpret_cast<uintptr_t>(data) + (size_t)base * esz) % 16 == 0) &&
                     (!flags || (reinterpret_cast<uintptr_t>(flags) + (size_t)base) % 4 == 0) &&
                     (!truth || (reinterpret_cast<uintptr_t>(truth) + (size_t)base) % 4 == 0);

    // shift for the moment sums: the pair's first sample (a typical value unless it is an outlier, in
    // which case the conditioning check below sends the pair through the two-pass formula)
    float K = 0.f;
    if (seg > 0) {
        if constexpr (DT == RFI_C64) { const float2 z = __ldg(static_cast<const float2*>(data) + base); K = cabs_np<float>(z.x, z.y); }
        else K = __ldg(static_cast<const float*>(data) + base);
        if (!(fabsf(K) < 3.0e38f)) K = 0.f;   // NaN / inf
    }
    // ---- load
    uint32_t fmask = 0;     // bit e: element e of this thread is flagged
    double s_all = 0.0, s_cln = 0.0, ss_all = 0.0, ss_cln = 0.0;
    uint32_t nflag = 0, tp = 0, fp = 0, fn = 0, nan_all = 0, nan_cln = 0, mx_all = 0, mx_cln = 0;
    uint32_t mn_all = kPExcl, mn_cln = kPExcl;
#pragma unroll 2
    for (int g = 0; g < G; ++g) {
        const long long i0 = ((long long)g * NT + tid) * 4;
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t f4 = 0, t4 = 0;
        int nin = 0;
        if (vec && i0 + 4 <= seg) {
            load4_mag_fast<DT>(data, (size_t)(base + i0), q);
            if (flags) f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + base + i0));
            if (truth) t4 = __ldg(reinterpret_cast<const uint32_t*>(truth + base + i0));
            nin = 4;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i0 + j < seg) {
                    if constexpr (DT == RFI_C64) {
                        const float2 z = __ldg(static_cast<const float2*>(data) + base + i0 + j);
                        q[j] = cabs_np<float>(z.x, z.y);
                    } else {
                        q[j] = __ldg(static_cast<const float*>(data) + base + i0 + j);
                    }
                    if (flags) f4 |= (uint32_t)__ldg(flags + base + i0 + j) << (8 * j);
                    if (truth) t4 |= (uint32_t)__ldg(truth + base + i0 + j) << (8 * j);
                    nin = j + 1;
                }
            }
        }
        // high bit of every byte = byte != 0
        const uint32_t pz = (((f4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | f4) & 0x80808080u;
        const uint32_t tz = (((t4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t4) & 0x80808080u;
        tp += __popc(pz & tz); fp += __popc(pz & ~tz); fn += __popc(~pz & tz);
        nflag += __popc(pz);
        const uint32_t fb = ((pz >> 7) & 1u) | ((pz >> 14) & 2u) | ((pz >> 21) & 4u) | ((pz >> 28) & 8u);
        fmask |= fb << (4 * g);
        PK k4[4];
        float sa = 0.f, sc = 0.f, qa = 0.f, qc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool in = j < nin, fl = (fb >> j) & 1u;
            const PK k = to_key<float>(q[j]);
            k4[j] = in ? k : kPExcl;
            const bool nan = in && k == kPExcl;
            nan_all += nan ? 1u : 0u;
            nan_cln += (nan && !fl) ? 1u : 0u;
            const PK km = in ? k : PK(0);                 // np.max propagates NaN: the all-ones key
            mx_all = km > mx_all ? km : mx_all;
            const PK kc = fl ? PK(0) : km;
            mx_cln = kc > mx_cln ? kc : mx_cln;
            mn_all = k4[j] < mn_all ? k4[j] : mn_all;
            const PK kd = fl ? kPExcl : k4[j];
            mn_cln = kd < mn_cln ? kd : mn_cln;
            const float d = q[j] - K;
            const float x = in ? d : 0.f, xc = (in && !fl) ? d : 0.f;
            sa += x; sc += xc;
            qa += x * x; qc += xc * xc;
        }
        s_all += (double)sa; s_cln += (double)sc;
        ss_all += (double)qa; ss_cln += (double)qc;
        *reinterpret_cast<uint4*>(skeys + ((size_t)g * NT + tid) * 4) = make_uint4(k4[0], k4[1], k4[2], k4[3]);
    }
    mx_all = warp_max(mx_all);
    mx_cln = warp_max(mx_cln);
    mn_all = warp_min(mn_all);
    mn_cln = warp_min(mn_cln);
    if (tid == 0) { sh.nxt[0] = 0; sh.nxt[1] = 0; sh.ext[0] = kPExcl; sh.ext[1] = kPExcl; }
    __syncthreads();
    if ((tid & 31) == 0) {
        atomicMax(&sh.nxt[0], mx_all); atomicMax(&sh.nxt[1], mx_cln);
        atomicMin(&sh.ext[0], mn_all); atomicMin(&sh.ext[1], mn_cln);
    }
    double d2[4] = {s_all, s_cln, ss_all, ss_cln};
    uint32_t u6[6] = {nflag, tp, fp, fn, nan_all, nan_cln};
    pair_totals<4, 6>(d2, u6, sh);   // (its barriers also publish the extremes)
    mx_all = sh.nxt[0]; mx_cln = sh.nxt[1];
    mn_all = sh.ext[0]; mn_cln = sh.ext[1];
    nflag = u6[0]; tp = u6[1]; fp = u6[2]; fn = u6[3]; nan_all = u6[4]; nan_cln = u6[5];
    const uint32_t n_cln = n_all - nflag;
    // the load pass summed d = x - K: sum x = sum d + n K
    const float mean_all = n_all ? (float)((d2[0] + (double)n_all * (double)K) / (double)n_all) : 0.f;
    const float mean_cln = n_cln ? (float)((d2[1] + (double)n_cln * (double)K) / (double)n_cln) : 0.f;

    // ---- sum of squared deviations from the T-rounded mean m (np.std: abs(x - m) ** 2, summed), with
    //      e = m - K:  sum (x - m)^2 = sum d^2 - 2 e sum d + n e^2, from the float64 sums of the load pass.
    //      The squares were rounded to float32 (~1e-9 of sum d^2 after averaging), so the result is good to
    //      1e-7 as long as it is not a small difference: it must exceed 1 % of sum d^2 (K within a few sigma
    //      of the mean); otherwise the data is swept again with the differences formed one by one
    double dq[2];
    bool resweep = false;
    {
        const double ea = (double)mean_all - (double)K, ec = (double)mean_cln - (double)K;
        dq[0] = d2[2] - 2.0 * ea * d2[0] + (double)n_all * ea * ea;
        dq[1] = d2[3] - 2.0 * ec * d2[1] + (double)n_cln * ec * ec;
        resweep = !(dq[0] > 0.01 * d2[2]) || (n_cln && !(dq[1] > 0.01 * d2[3]));   // also true for NaN / inf sums
    }
    if (resweep) {
    double q_all = 0.0, q_cln = 0.0;
#pragma unroll 2
    for (int g = 0; g < G; ++g) {
        const uint4 kq = pair_keys(skeys, g);
        const PK k4[4] = {kq.x, kq.y, kq.z, kq.w};
        const long long i0 = ((long long)g * NT + tid) * 4;
        float qa = 0.f, qc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool in = i0 + j < seg, fl = (fmask >> (4 * g + j)) & 1u;
            const float x = from_key<float>(k4[j]);   // the all-ones key decodes to NaN
            const float da = x - mean_all, dc = x - mean_cln;
            qa += in ? da * da : 0.f;
            qc += (in && !fl) ? dc * dc : 0.f;
        }
        q_all += (double)qa;
        q_cln += (double)qc;
    }
    dq[0] = q_all; dq[1] = q_cln;
    uint32_t u0[1] = {0};
    pair_totals<2, 1>(dq, u0, sh);
    }

    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    rfi_stats_t oa, oc;
    oa.count = (long long)n_all; oa.n_flagged = 0; oa.n_nan = (long long)nan_all;
    oc.count = (long long)n_cln; oc.n_flagged = (long long)nflag; oc.n_nan = (long long)nan_cln;
    oa.mean = oa.std = oa.median = oa.mad = oa.max = kNaN;
    oc.mean = oc.std = oc.median = oc.mad = oc.max = kNaN;
    if (n_all) {
        oa.mean = (double)mean_all;
        oa.std = (double)__fsqrt_rn((float)(dq[0] / (double)n_all));
        oa.max = (double)from_key<float>(mx_all);
    }
    if (n_cln) {
        oc.mean = (double)mean_cln;
        oc.std = (double)__fsqrt_rn((float)(dq[1] / (double)n_cln));
        oc.max = (double)from_key<float>(mx_cln);
    }

    // ---- order statistics: all samples, then the unflagged ones
    constexpr PK kInfHi = to_key_const_inf<float>(false), kInfLo = to_key_const_inf<float>(true);
    if (n_all && nan_all == 0) {
        const bool any_inf = mx_all >= kInfHi || mn_all <= kInfLo;   // +-inf present: radix route
        float med, mad;
        pair_median_mad(skeys, cand, samp, sh, n_all, any_inf, med, mad);
        oa.median = (double)med; oa.mad = (double)mad;
        if (nflag == 0) { oc.median = oa.median; oc.mad = oa.mad; }
    }
    if (nflag != 0 && n_cln && nan_cln == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint32_t fb = (fmask >> (4 * g)) & 15u;
            if (fb) {
                PK* kp = skeys + ((size_t)g * NT + tid) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((fb >> j) & 1u) kp[j] = kPExcl;
            }
        }
        __syncthreads();
        const bool any_inf = mx_cln >= kInfHi || mn_cln <= kInfLo;
        float med, mad;
        pair_median_mad(skeys, cand, samp, sh, n_cln, any_inf, med, mad);
        oc.median = (double)med; oc.mad = (double)mad;
    } else if (nflag == 0 && n_all && nan_all != 0) {
        // unflagged == all, NaN among them: stays NaN
    }

    if (tid == 0) {
        if (out) { out[(long long)blockIdx.x * 2] = oa; out[(long long)blockIdx.x * 2 + 1] = oc; }
        if (res) {
            // statistics.py:73-97 on Python floats: float64, one IEEE operation per source operation
            rfi_pair_result_t r;
            r.tp = tp; r.fp = fp; r.fn = fn;
            {   // metrics.py:39-45, 66-79, 98-104, 120-126, 145-152 on the integer counts, in float64
                const double dtp = (double)tp, dfp = (double)fp, dfn = (double)fn;
                const uint32_t uni = tp + fp + fn;
                r.iou = uni == 0 ? 1.0 : dtp / (double)uni;
                r.precision = (tp + fp == 0) ? (fn == 0 ? 1.0 : 0.0) : dtp / (double)(tp + fp);
                r.recall = (tp + fn == 0) ? 1.0 : dtp / (double)(tp + fn);
                const double pr = r.precision + r.recall;
                r.f1 = pr == 0.0 ? 0.0 : 2.0 * (r.precision * r.recall) / pr;
                r.dice = (2 * tp + fp + fn == 0) ? 1.0 : (double)(2 * tp) / (double)(2 * tp + fp + fn);
                (void)dfp; (void)dfn;
            }
            const double frac = n_all ? (double)nflag / (double)n_all : kNaN;   // np.sum(flags) / flags.size
            const bool guard = n_cln == 0 || oc.mad != oc.mad || oc.std != oc.std;  // :77-78
            if (guard) {
                r.ffi = r.mad_reduction = r.std_reduction = 0.0; r.flagged_fraction = 1.0; r.status = 1;
            } else if (oa.mad == 0.0 || oa.std == 0.0) {   // the reference raises ZeroDivisionError
                r.ffi = r.mad_reduction = r.std_reduction = kNaN; r.flagged_fraction = frac; r.status = 2;
            } else {
                const double mad_red = 1.0 - (oc.mad / oa.mad);
                const double std_red = 1.0 - (oc.std / oa.std);
                const double a = 0.5 * mad_red, b = 0.5 * std_red, pen = 0.5 * frac;
                r.mad_reduction = mad_red; r.std_reduction = std_red; r.flagged_fraction = frac;
                r.ffi = (a + b) * (1.0 - pen);
                r.status = 0;
            }
            res[blockIdx.x] = r;
        }
    }
}

}  // namespace rfi

using namespace rfi;

extern "C" int rfi_pair_sweep(const void* data, int dtype, const uint8_t* flags, const uint8_t* truth,
                              int64_t n_pairs, int64_t seg, rfi_stats_t* stats, rfi_pair_result_t* results,
                              void* stream) {
    if (n_pairs < 0 || seg <= 0 || (n_pairs > 0 && !data)) { set_error("bad arguments to rfi_pair_sweep"); return RFI_E_INVALID; }
    if (dtype != RFI_F32 && dtype != RFI_C64) { set_error("rfi_pair_sweep: float32 / complex64 data only (dtype %d)", dtype); return RFI_E_UNSUPPORTED; }
    if (seg > kPairSeg) { set_error("rfi_pair_sweep: pairs of at most %d samples (got %lld)", kPairSeg, (long long)seg); return RFI_E_UNSUPPORTED; }
    if (n_pairs == 0) return RFI_OK;
    if (n_pairs > 0x7fffffffLL) { set_error("too many pairs for one launch"); return RFI_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(kPairSeg + kMonoCap + kPairNT) * sizeof(PK);
    if (dtype == RFI_F32) {
        auto k = pair_sweep_kernel<RFI_F32>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(unsigned)n_pairs, kPairNT, smem, st>>>(data, flags, truth, seg, stats, results);
    } else {
        auto k = pair_sweep_kernel<RFI_C64>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(unsigned)n_pairs, kPairNT, smem, st>>>(data, flags, truth, seg, stats, results);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
