// rfi_select_keys.cuh -- exact median / MAD of up to 16384 order-preserving float32 keys held by one CTA of
// 512 threads, 32 keys per thread in eight groups of four, reached through an accessor (shared memory, or
// partly a thread-private global scratch): the sampled-bracket selection of phase 1 (rfi_stats_mono.cuh)
// for ARBITRARY keys -- no monotone chain assumed -- with an MSB-first radix select over the same keys as
// the fallback.  Used by the per-pair sweep (rfi_pairs.cu).  (Tried for the general tile statistics of
// create_dataset too -- DESIGN.md 9.2: any change to the 40-register statistics kernel cost its main path
// more than the rare tiles gained.)
//
// Accessor KA:  uint4 load(int g) const   -- this thread's four keys of group g (0 .. 7)
//               PK load1(int g, int i) const -- key i of group g
// Excluded samples (NaN, flagged, padding) carry the all-ones key and are not counted in nv.
#pragma once
#include "rfi_stats_mono.cuh"

namespace rfi {

constexpr int kSelNT = 512, kSelG = 8;
using PK = uint32_t;
constexpr PK kSelExcl = ~PK(0);

struct SelShared {
    MonoShared<PK> ms;
    uint32_t cnt[4][16];  // radix-select counters, rotating (see RoundCounter)
    uint32_t nxt[2];      // count(key <= prefix), min(key > prefix)
};

// ---- exact fallback: MSB-first radix select over the shared-memory keys, 4 bits per pass, 15
// register counters per thread (no histogram atomics).  DEV: the key of a sample is its absolute
// deviation from `centre`.  Returns the keys of ranks k1 <= k2 <= k1 + 1 among the valid keys.
template <bool DEV, typename KA>
__device__ __noinline__ void sel_radix(const KA ka, uint32_t k1, uint32_t k2, float centre,
                                               PK& o1, PK& o2, SelShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31;
    auto key_of = [&](PK x) -> PK {
        if (!DEV) return x;
        return x == kSelExcl ? kSelExcl : to_key<float>(fabsf(from_key<float>(x) - centre));
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
        for (int g = 0; g < kSelG; ++g) {
            const uint4 q = ka.load(g);
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
        PK nxt = kSelExcl;
#pragma unroll 2
        for (int g = 0; g < kSelG; ++g) {
            const uint4 q = ka.load(g);
            const PK k4[4] = {key_of(q.x), key_of(q.y), key_of(q.z), key_of(q.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cle += (k4[i] <= prefix) ? 1u : 0u;
                const PK y = k4[i] > prefix ? k4[i] : kSelExcl;
                nxt = y < nxt ? y : nxt;
            }
        }
        cle = __reduce_add_sync(0xffffffffu, cle);
        nxt = warp_min(nxt);
        __syncthreads();
        if (tid == 0) { sh.nxt[0] = 0; sh.nxt[1] = kSelExcl; }
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
template <typename KA>
RFI_DEVINL int sel_sort_sample(const KA ka, PK* runs, PK* samp) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int e_s = (((lane + warp * 3) & 7) << 2) | (((lane >> 3) + warp) & 3);
    PK x = ka.load1(e_s >> 2, e_s & 3);
    runs[tid] = x;
    const int nvalid = __syncthreads_count(x != kSelExcl);
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

// the thread's keys whose bit is set in `marked` (bit e = key i of group g, e = g * 4 + i), appended to
// cand[at ...] in ascending element order; with HIST each is also counted into sh.hist (bucket
// (key - hlo) >> hshf): mono_resolve's first histogram, built on the way
template <bool HIST, typename KA>
RFI_DEVINL void sel_gather_marked(const KA ka, uint32_t marked, PK* cand, uint32_t at, MonoShared<PK>& sh, PK hlo, int hshf) {
    while (marked) {
        const int e = __ffs((int)marked) - 1;
        marked &= marked - 1u;
        const PK x = ka.load1(e >> 2, e & 3);
        cand[at++] = x;
        if (HIST) atomicAdd(&sh.hist[(uint32_t)((PK)(x - hlo) >> hshf)], 1u);
    }
}

// ---- two middle order statistics of the nv valid keys by a sampled bracket (one directional retry).
// false = bracket missed / too many candidates: the caller runs the radix select.
template <typename KA>
__device__ __noinline__ bool sel_sampled_median(const KA ka, PK* cand, const PK* samp,
                                                 SelShared& shp, uint32_t nv, int sv, PK& v1k, PK& v2k) {
    MonoShared<PK>& sh = shp.ms;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    const int delta = (int)(kMonoSigma * 0.5f * sqrtf((float)sv)) + 2;
    const int rho = (int)(((float)(2 * k1 + 1) * (float)sv) / (float)(2 * nv));
    const int ilo = rho - delta, ihi = rho + delta + 1;
    PK lo = ilo < 0 ? PK(0) : samp[ilo];
    PK hi = ihi >= sv ? kSelExcl - 1 : samp[ihi];
    if (tid == 0) { sh.cursor = 0; sh.below = 0; }
    __syncthreads();
    uint32_t M = 0, B = 0;
#pragma unroll 1
    for (int attempt = 0;; ++attempt) {
        const PK span = hi - lo;
        for (int b = tid; b < kMonoBuckets; b += kSelNT) sh.hist[b] = 0;   // filled by the compaction (barrier below)
        uint32_t below = 0, inmask = 0;   // one bit per key inside the bracket: the compaction visits only those
#pragma unroll
        for (int g = 0; g < kSelG; ++g) {
            const uint4 q = ka.load(g);
            const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                below += (k4[i] < lo) ? 1u : 0u;
                if ((PK)(k4[i] - lo) <= span) inmask |= 1u << (g * 4 + i);
            }
        }
        const uint32_t mine = __popc(inmask);
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
            const bool dead = attempt == 1 || M > (uint32_t)kMonoCap || (low && lo == 0) || (high && hi >= kSelExcl - 1);
            __syncthreads();  // every thread has read the totals
            if (dead) return false;
            if (low) {
                const int j = ilo - 2 * delta;
                hi = lo - 1;
                lo = j < 0 ? PK(0) : samp[j];
            } else {
                const int j = ihi + 2 * delta;
                lo = hi + 1;
                hi = j >= sv ? kSelExcl - 1 : samp[j];
            }
            if (tid == 0) { sh.cursor = 0; sh.below = 0; }
            __syncthreads();
            continue;
        }
        sel_gather_marked<true>(ka, inmask, cand, at, sh, lo, mono_bucket_shift<PK>(span));
        break;
    }
    __syncthreads();
    mono_resolve<PK, kSelNT>(cand, M, k1 - B, k2 - B, v1k, v2k, sh, true, lo, hi);
    return true;
}

// ---- two middle order statistics of |x - c| over the valid keys: |x - c| is V-shaped in x, so the
// r smallest deviations are a contiguous window of the sorted sample around c; an inner and an outer
// window prove bounds for everything strictly inside / outside, only the candidates in between get
// their exact deviation, and the answer is accepted only inside what was proven.
template <typename KA>
__device__ __noinline__ bool sel_sampled_mad(const KA ka, PK* cand, const PK* samp,
                                              SelShared& shp, uint32_t nv, int sv, float c, PK& r1k, PK& r2k) {
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
    const PK U2 = iu2 + 1 < sv ? samp[iu2 + 1] : kSelExcl - 1;
    const PK d_in = dsamp[il] > dsamp[iu] ? dsamp[il] : dsamp[iu];       // interior deviations <= this
    const PK d_lo2 = il2 > 0 ? dsamp[il2 - 1] : kSelExcl, d_up2 = iu2 + 1 < sv ? dsamp[iu2 + 1] : kSelExcl;
    const PK d_out = d_lo2 < d_up2 ? d_lo2 : d_up2;                      // exterior deviations >= this
    // every candidate deviates at least / at most this much (each arm of the V is monotone): the range of the
    // histogram that is filled while the deviations are formed
    const PK d_min = dsamp[il] < dsamp[iu] ? dsamp[il] : dsamp[iu];
    const PK d_max = d_lo2 > d_up2 ? d_lo2 : d_up2;
    __syncthreads();  // dsamp (= cand) is overwritten below
    for (int b = tid; b < kMonoBuckets; b += kSelNT) sh.hist[b] = 0;
    const PK span_all = U2 - L2;
    const PK w_in = U1 > L1 ? U1 - L1 - 1 : PK(0);
    uint32_t inside = 0, inmask = 0;
#pragma unroll
    for (int g = 0; g < kSelG; ++g) {
        const uint4 q = ka.load(g);
        const PK k4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool in_all = (PK)(k4[i] - L2) <= span_all;
            const bool interior = (PK)(k4[i] - L1 - 1) < w_in;
            inside += interior ? 1u : 0u;
            if (in_all && !interior) inmask |= 1u << (g * 4 + i);
        }
    }
    const uint32_t mine = __popc(inmask);
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
    sel_gather_marked<false>(ka, inmask, cand, at, sh, PK(0), 0);
    __syncthreads();
    const PK d_width = d_max - d_min;
    const int d_shf = mono_bucket_shift<PK>(d_width);
    bool outside = false;
    for (uint32_t i = tid; i < M; i += kSelNT) {
        const PK dk = to_key<float>(fabsf(from_key<float>(cand[i]) - c));
        cand[i] = dk;
        if ((PK)(dk - d_min) <= d_width) atomicAdd(&sh.hist[(uint32_t)((PK)(dk - d_min) >> d_shf)], 1u);
        else outside = true;
    }
    const bool hist_ok = __syncthreads_or(outside) == 0;   // (else the list is scanned for its range and histogram)
    mono_resolve<PK, kSelNT>(cand, M, k1 - B, k2 - B, r1k, r2k, sh, hist_ok, d_min, d_max);
    return !(r1k < d_in || r2k > d_out);  // the answers must lie inside what the windows prove
}

// two middle order statistics of the nv valid keys: sampled bracket, radix select if it misses.
// `sv` (out): valid samples of the sorted sample left in `samp` (0 = no sample was drawn), for sel_mad.
template <typename KA>
__device__ __noinline__ void sel_median(const KA ka, PK* cand, PK* samp, SelShared& sh, uint32_t nv, bool sample_ok,
                                        PK& a, PK& b, int& sv) {
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    bool sampled = sample_ok && nv >= 64;
    sv = 0;
    if (sampled) {
        sv = sel_sort_sample(ka, cand, samp);
        sampled = sv >= 64;
        if (!sampled) sv = 0;
    }
    if (!(sampled && sel_sampled_median(ka, cand, samp, sh, nv, sv, a, b)))
        sel_radix<false>(ka, k1, k2, 0.f, a, b, sh);
}

// two middle order statistics of |x - c| over the nv valid keys (`samp`, `sv` from sel_median on the SAME keys)
template <typename KA>
__device__ __noinline__ void sel_mad(const KA ka, PK* cand, const PK* samp, SelShared& sh, uint32_t nv, int sv, float c,
                                     PK& a, PK& b) {
    const uint32_t k1 = (nv - 1) >> 1, k2 = nv >> 1;
    if (!(sv >= 64 && sel_sampled_mad(ka, cand, samp, sh, nv, sv, c, a, b)))
        sel_radix<true>(ka, k1, k2, c, a, b, sh);
}

}  // namespace rfi
