// rfi_bigtile.cu -- create_dataset for the large legal patch sizes (P = 256 / 512 / 1024,
// config/validators.py:26-39) when the waterfall dims are multiples of P: float32 arithmetic
// (float32 / complex64 input), real branch.  BASELINE config 5 (P = 256) runs here.
//
// A P x P group (65 536 .. 1 Mi samples per median) no longer fits one CTA's shared memory, so
// the sampled-bracket selection of the P = 128 kernel (rfi_stats_mono.cuh) is split into
// stream-ordered launches over 128 x 128 SUB-TILES (one CTA each, same thread / sample mapping
// as the P = 128 kernels) and over GROUPS (one CTA each), with the group state in global memory:
//
//   big_load      sub-tile  |z| fused into the 128-bit loads, exact magnitudes written ONCE to a
//                           float32 scratch (every later pass, phase 2 included, reads 4 B / px
//                           instead of 8 and never repeats the division / square root);
//                           raw extremes; stratified 2048-sample of the group
//   big_sample    group     sorts the sample (bitonic, shared memory), median bracket (4.5 sigma)
//   big_pass<0>   sub-tile  counts the keys below the bracket, compacts the ones inside
//   big_median    group     exact ranks inside the candidates (mono_resolve), normalisation
//                           medians, centre; V-shaped deviation windows of the sorted sample
//   big_pass<1>   sub-tile  counts the interior, compacts the two candidate rings
//   big_mad       group     exact deviations of the candidates, MAD, thresholds, raw thresholds
//   (host)        groups whose brackets missed or that hold NaN / inf / negative / inf-filled
//                 samples are listed; the generic select (rfi_generic.cu) measures those
//   big_range     sub-tile  flag counts; log-amplitude tile + halo rows / columns -> per-group
//                           min / max of L and of the three squared-gradient variants
//   big_write     sub-tile  phase 2: every kept rotation of the sub-tile written into its block
//                           of the P x P output patch, gradients continuous across sub-tiles
//
// Every bracket is validated exactly as in the P = 128 kernel, so results are exact either way.
// Reference semantics: preprocessor.py:22-42, 413-446, 562-783.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <vector>

#include "rfi_tiles.cuh"
#include "rfi_stats_mono.cuh"

namespace rfi {

constexpr int kBigS = 2048;   // sorted sample per group
constexpr int kBigNT = 512;   // threads of the sub-tile kernels (one 128 x 128 sub-tile per CTA)

struct BigGroup {             // per-group state (workspace)
    uint32_t kmin, kmax;      // raw extremes (bit patterns)
    uint32_t cursor, below;   // candidates appended so far; keys below the bracket / inside the inner window
    uint32_t lo, hi;          // median bracket (inclusive keys)
    uint32_t L1, U1, L2, U2;  // MAD: inner window (L1, U1) and outer ring [L2, U2]
    uint32_t d_in, d_out;     // proven bounds of the deviations inside / outside
    float m, m2, c;           // median before, median after, centre
    int fail;                 // != 0: measured by the generic select instead (reason as in the P = 128 kernel)
    uint32_t rng[8];          // ordered keys: min L, max L, then (min, max) of the squared gradient variants 0, 1, 3
};

struct BigGeom {
    PlanDev p;                // nh / nw = GROUP tiles per waterfall
    int P, n, n2;             // patch size, sub-tiles per side / per group
    long long n_groups;
    uint32_t cap;             // candidate capacity per group
    int delta;                // half width of the sample-rank brackets
};

struct BigTile {              // one sub-tile (uniform per CTA)
    long long grp, w;
    int TI, TJ, si, sj;
    size_t origin;            // first sample of the sub-tile inside the cube
    bool hasT, hasB, hasL, hasR;  // neighbouring sub-tile inside the same group
};

RFI_DEVINL BigTile big_tile(const BigGeom& g, long long blk) {
    BigTile t;
    t.grp = blk / g.n2;
    const int sub = (int)(blk % g.n2);
    t.si = sub / g.n; t.sj = sub % g.n;
    const int per = g.p.nh * g.p.nw;
    t.w = t.grp / per;
    const int r = (int)(t.grp % per);
    t.TI = r / g.p.nw; t.TJ = r % g.p.nw;
    t.origin = ((size_t)t.w * g.p.channels + (size_t)t.TI * g.P + (size_t)t.si * kP) * g.p.times +
               (size_t)t.TJ * g.P + (size_t)t.sj * kP;
    t.hasT = t.si > 0; t.hasB = t.si < g.n - 1; t.hasL = t.sj > 0; t.hasR = t.sj < g.n - 1;
    return t;
}

RFI_DEVINL void big_fail(BigGroup* groups, long long grp, int reason, int* fail_list, int* fail_count) {
    if (threadIdx.x == 0) {
        groups[grp].fail = reason;
        fail_list[atomicAdd(fail_count, 1)] = (int)grp;
    }
}

// ------------------------------------------------------------------------------------------
__global__ void big_init_kernel(BigGeom g, BigGroup* __restrict__ groups, int* __restrict__ fail_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *fail_count = 0;
    if (i >= g.n_groups) return;
    BigGroup G;
    G.kmin = ~0u; G.kmax = 0; G.cursor = 0; G.below = 0; G.lo = G.hi = 0;
    G.L1 = G.U1 = G.L2 = G.U2 = 0; G.d_in = G.d_out = 0; G.m = G.m2 = G.c = 0.f; G.fail = 0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) { G.rng[k] = ~0u; G.rng[k + 1] = 0u; }
    groups[i] = G;
}

// re-arms the range accumulators of the listed groups (after the generic fallback measured them)
__global__ void big_rearm_kernel(BigGroup* __restrict__ groups, const int* __restrict__ list, int n_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    BigGroup& G = groups[list[i]];
#pragma unroll
    for (int k = 0; k < 8; k += 2) { G.rng[k] = ~0u; G.rng[k + 1] = 0u; }
}

// ------------------------------------------------------------------------------------------
// sub-tile load: magnitude scratch, extremes, sample
template <int DT>
__global__ void __launch_bounds__(kBigNT, 2)
big_load_kernel(BigGeom g, const void* __restrict__ data, float* __restrict__ mag,
                BigGroup* __restrict__ groups, uint32_t* __restrict__ samples, int want_sample) {
    constexpr int G8 = kP * kP / kBigNT / 4, RS = kBigNT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BigTile t = big_tile(g, blockIdx.x);
    const int e_s = (((lane + warp * 3) & 7) << 2) | (((lane >> 3) + warp) & 3);  // stratified, as the P = 128 kernel
    uint32_t bmax = 0, bmin = ~0u, mine = 0;
#pragma unroll
    for (int g8 = 0; g8 < G8; ++g8) {
        const size_t idx = t.origin + (size_t)(g8 * RS + warp) * g.p.times + lane * 4;
        float q[4];
        load4_mag_fast<DT>(data, idx, q);
        if constexpr (In<DT>::cplx) *reinterpret_cast<float4*>(mag + idx) = make_float4(q[0], q[1], q[2], q[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t k = __float_as_uint(q[i]);
            bmax = k > bmax ? k : bmax;
            bmin = k < bmin ? k : bmin;
            if (g8 * 4 + i == e_s) mine = k;
        }
    }
    if (!want_sample) return;
    __shared__ uint32_t smin, smax;
    if (tid == 0) { smin = ~0u; smax = 0; }
    __syncthreads();
    bmax = warp_max(bmax);
    bmin = warp_min(bmin);
    if (lane == 0) { atomicMin(&smin, bmin); atomicMax(&smax, bmax); }
    __syncthreads();
    if (tid == 0) { atomicMin(&groups[t.grp].kmin, smin); atomicMax(&groups[t.grp].kmax, smax); }
    // kBigS / n2 samples per sub-tile: thread tid samples iff tid % stride == (tid / stride) % stride
    const int spc = kBigS / g.n2, stride = kBigNT / spc;
    const int slot = tid / stride;
    if (tid % stride == slot % stride)
        samples[(size_t)t.grp * kBigS + (size_t)(t.si * g.n + t.sj) * spc + slot] = mine;
}

// ------------------------------------------------------------------------------------------
// group: sort the sample, median bracket.  Four warps each sort 512 keys in registers (bitonic
// network over 16 keys per lane, shuffles across lanes -- no shared-memory traffic, no block
// barriers), then every key finds its final rank by binary searches in the other three runs
// (ties broken by run index, so the ranks are a permutation).
__global__ void __launch_bounds__(128)
big_sample_kernel(BigGeom g, BigGroup* __restrict__ groups, uint32_t* __restrict__ samples,
                  int* __restrict__ fail_list, int* __restrict__ fail_count) {
    constexpr int kRun = 512, kRuns = kBigS / kRun;
    static_assert(kRuns == 4, "one run per warp");
    __shared__ uint32_t runs[kBigS], s[kBigS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long grp = blockIdx.x;
    uint32_t* gs = samples + (size_t)grp * kBigS;
    uint32_t v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = gs[warp * kRun + r * 32 + lane];
    warp_sort512<uint32_t>(v, lane);
#pragma unroll
    for (int r = 0; r < 16; ++r) runs[warp * kRun + r * 32 + lane] = v[r];
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < 16; ++r) {
        uint32_t rank = r * 32 + lane;
        const uint32_t x = runs[warp * kRun + rank];  // = v[r], without indexing the register array
#pragma unroll
        for (int o = 0; o < kRuns; ++o) {
            if (o == warp) continue;  // warp-uniform
            const uint32_t* run = runs + o * kRun;
            const bool incl = o < warp;  // earlier runs win ties
            uint32_t pos = 0;
#pragma unroll
            for (int step = kRun / 2; step > 0; step >>= 1) {
                const uint32_t y = run[pos + step - 1];
                pos += (incl ? (y <= x) : (y < x)) ? step : 0;
            }
            const uint32_t y = run[kRun - 1];
            pos += (pos == kRun - 1 && (incl ? (y <= x) : (y < x))) ? 1u : 0u;
            rank += pos;
        }
        s[rank] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBigS / 128; ++k) gs[tid + k * 128] = s[tid + k * 128];
    if (tid == 0) {
        BigGroup& G = groups[grp];
        if (G.kmax >= 0x7f800000u) {   // NaN, inf or a negative sample
            G.fail = 1;
            fail_list[atomicAdd(fail_count, 1)] = (int)grp;
        } else {
            const uint32_t nv = (uint32_t)g.P * (uint32_t)g.P, k1 = (nv - 1) >> 1;
            const int rho = (int)(((double)(2 * k1 + 1) * (double)kBigS) / (double)(2.0 * nv));
            const int ilo = rho - g.delta, ihi = rho + g.delta + 1;
            G.lo = ilo < 0 ? 0u : s[ilo];
            G.hi = ihi >= kBigS ? 0xfffffffeu : s[ihi];
        }
    }
}

// ------------------------------------------------------------------------------------------
// sub-tile pass over the raw keys: count + compact.  MODE 0: median bracket; MODE 1: MAD rings.
template <int MODE>
__global__ void __launch_bounds__(kBigNT, 3)
big_pass_kernel(BigGeom g, const float* __restrict__ src, BigGroup* __restrict__ groups,
                uint32_t* __restrict__ cand_all) {
    constexpr int G8 = kP * kP / kBigNT / 4, RS = kBigNT / 32, W = kBigNT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BigTile t = big_tile(g, blockIdx.x);
    BigGroup* G = groups + t.grp;
    if (G->fail) return;
    uint32_t a0, span0, a1 = 0, span1 = 0;   // mine: (k - a0) <= span0 [and not (k - a1) < span1]; count: below a0 / inside (a1, span1)
    if (MODE == 0) { a0 = G->lo; span0 = G->hi - G->lo; }
    else { a0 = G->L2; span0 = G->U2 - G->L2; a1 = G->L1 + 1; span1 = G->U1 > G->L1 ? G->U1 - G->L1 - 1 : 0u; }
    auto is_mine = [&](uint32_t k) {
        if (MODE == 0) return (uint32_t)(k - a0) <= span0;
        return ((uint32_t)(k - a0) <= span0) && !((uint32_t)(k - a1) < span1);
    };
    auto is_counted = [&](uint32_t k) {
        if (MODE == 0) return k < a0;
        return (uint32_t)(k - a1) < span1;
    };
    // sweep 1: counts only.  The keys are not kept in registers (three CTAs per SM hide the
    // round trip of the group's cursor atomic); sweep 2 reads them again from L1 / L2.
    uint32_t cnt = 0, mine = 0;
    const float* row0 = src + t.origin + (size_t)warp * g.p.times + lane * 4;
    const size_t pitch = (size_t)RS * g.p.times;
    {
        float4 q[G8];
#pragma unroll
        for (int g8 = 0; g8 < G8; ++g8) q[g8] = __ldg(reinterpret_cast<const float4*>(row0 + g8 * pitch));
#pragma unroll
        for (int g8 = 0; g8 < G8; ++g8) {
            const uint32_t k4[4] = {__float_as_uint(q[g8].x), __float_as_uint(q[g8].y), __float_as_uint(q[g8].z), __float_as_uint(q[g8].w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) { cnt += is_counted(k4[i]) ? 1u : 0u; mine += is_mine(k4[i]) ? 1u : 0u; }
        }
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ uint32_t wtot[W], wcnt[W], base_s, total_s;
    if (lane == 31) { wtot[warp] = incl; wcnt[warp] = cnt; }
    __syncthreads();
    if (tid == 0) {
        uint32_t tot = 0, c = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) { const uint32_t v = wtot[i]; wtot[i] = tot; tot += v; c += wcnt[i]; }
        base_s = tot ? atomicAdd(&G->cursor, tot) : 0u;
        total_s = tot;
        if (c) atomicAdd(&G->below, c);
    }
    __syncthreads();
    if (base_s + total_s > g.cap) return;  // overflow: the group kernel sees cursor > cap and gives up
    if (mine == 0) return;
    uint32_t* cand = cand_all + (size_t)t.grp * g.cap + base_s + wtot[warp] + (incl - mine);
#pragma unroll 2
    for (int g8 = 0; g8 < G8; ++g8) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(row0 + g8 * pitch));
        const uint32_t k4[4] = {__float_as_uint(q.x), __float_as_uint(q.y), __float_as_uint(q.z), __float_as_uint(q.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (is_mine(k4[i])) *cand++ = k4[i];
    }
}

// ------------------------------------------------------------------------------------------
// group: median from the candidates, processed-domain statistics, deviation windows
RFI_DEVINL void big_write_stat(rfi_tile_stat_t* out, const PlanDev& p, float m, float m2, float c, float d,
                               float thr_lo, float thr_hi, float raw_lo, float raw_hi, int nv) {
    rfi_tile_stat_t st;
    st.median_before = p.norm_before ? (double)m : 0.0;
    st.inf_fill = 0.0;
    st.median_after = p.norm_after ? (double)m2 : 0.0;
    st.centre = (double)c; st.mad = (double)d;
    st.thr_lo = (double)thr_lo; st.thr_hi = (double)thr_hi;
    st.n_valid = nv; st.n_inf = 0; st.n_flagged = 0;
    st.route = RFI_TILE_RAW_THRESHOLDS;
    st.raw_lo = (double)raw_lo; st.raw_hi = (double)raw_hi;
    *out = st;
}

__global__ void __launch_bounds__(kBigNT, 3)
big_median_kernel(BigGeom g, BigGroup* __restrict__ groups, const uint32_t* __restrict__ samples,
                  uint32_t* __restrict__ cand_all, rfi_tile_stat_t* __restrict__ stats,
                  int* __restrict__ fail_list, int* __restrict__ fail_count) {
    using T = float;
    using K = uint32_t;
    constexpr int NT = kBigNT;
    __shared__ MonoShared<K> sh;
    __shared__ K samp[kBigS], dsamp[kBigS];
    const int tid = threadIdx.x;
    const long long grp = blockIdx.x;
    const PlanDev& p = g.p;
    BigGroup* G = groups + grp;
    if (G->fail) return;
    const uint32_t nv = (uint32_t)g.P * (uint32_t)g.P, k1 = (nv - 1) >> 1, k2 = nv >> 1;
    const uint32_t M = G->cursor, B = G->below;
    const K tile_min = G->kmin, tile_max = G->kmax;
    if (M > g.cap || B > k1 || k2 >= B + M) { big_fail(groups, grp, 3, fail_list, fail_count); return; }
    K v1k, v2k;
    mono_resolve<K, NT>(cand_all + (size_t)grp * g.cap, M, k1 - B, k2 - B, v1k, v2k, sh);

    T m = T(0), m2 = T(0);
    const T v1 = raw_val<T>(v1k), v2 = raw_val<T>(v2k);
    if (p.norm_before) m = median_of_pair<T>(v1, v2, nv);
    T s1 = v1, s2 = v2;
    if (p.norm_before && m > T(0)) { s1 = s1 / m; s2 = s2 / m; }
    if (p.stretch != RFI_STRETCH_NONE) { s1 = apply_stretch<T>(s1, p.stretch); s2 = apply_stretch<T>(s2, p.stretch); }
    if (p.norm_after) {
        m2 = median_of_pair<T>(s1, s2, nv);
        if (m2 > T(0)) { s1 = s1 / m2; s2 = s2 / m2; }
    }
    {   // the extreme samples must stay finite through the chain (else: inf fill -> generic select)
        const T pmin = proc_nofill<T>(raw_val<T>(tile_min), p, m, m2);
        const T pmax = proc_nofill<T>(raw_val<T>(tile_max), p, m, m2);
        if (is_inf(pmin) || is_inf(pmax) || is_nan(pmin) || is_nan(pmax)) { big_fail(groups, grp, 5, fail_list, fail_count); return; }
    }
    if (p.flag_mode != RFI_FLAGS_MAD) {
        if (tid == 0) {
            G->m = m; G->m2 = m2;
            big_write_stat(stats + grp, p, m, m2, 0.f, 0.f, 0.f, 0.f, 0.f, Scalar<T>::inf(), (int)nv);
        }
        return;
    }
    const T c = median_of_pair<T>(s1, s2, nv);
    // ---- deviations of the sorted raw sample: V-shaped in the sample index
    constexpr int sv = kBigS;
    int ju = 0;
    if (tid == 0) { sh.win[0] = sh.win[1] = sh.win[2] = sh.win[3] = -1; }
#pragma unroll 1
    for (int i = tid; i < sv; i += NT) {
        const K sk = samples[(size_t)grp * kBigS + i];
        samp[i] = sk;
        const T ps = proc_nofill<T>(raw_val<T>(sk), p, m, m2);
        dsamp[i] = to_key<T>(fabs_(ps - c));
        ju += __syncthreads_count(ps < c);  // first sample on the upper arm (proc >= c)
    }
    __syncthreads();
    const int rho = (int)(((double)(2 * k1 + 1) * (double)sv) / (2.0 * (double)nv));
    const int r_in = rho - g.delta, r_out = rho + g.delta + 2;
    if (r_in < 1 || r_out > sv - 1) { big_fail(groups, grp, 6, fail_list, fail_count); return; }
    for (int which = 0; which < 2; ++which) {
        const int r = which == 0 ? r_in : r_out;
        auto pred = [&](int s) {
            if (s + r >= sv) return true;
            return dsamp[s] <= dsamp[s + r] && (s + r) >= ju;
        };
#pragma unroll 1
        for (int i = tid; i + r <= sv; i += NT)
            if (pred(i) && (i == 0 || !pred(i - 1))) { sh.win[which * 2] = i; sh.win[which * 2 + 1] = i + r - 1; }
    }
    __syncthreads();
    int il = sh.win[0], iu = sh.win[1], il2 = sh.win[2], iu2 = sh.win[3];
    if (il < 0 || il2 < 0) { big_fail(groups, grp, 7, fail_list, fail_count); return; }
    il2 = il2 < il ? il2 : il;
    iu2 = iu2 > iu ? iu2 : iu;
    if (!(il2 <= il && il < ju && ju <= iu && iu <= iu2 && il2 < ju)) { big_fail(groups, grp, 8, fail_list, fail_count); return; }
    if (tid == 0) {
        G->L1 = samp[il]; G->U1 = samp[iu];
        G->L2 = il2 > 0 ? samp[il2 - 1] : 0u;
        G->U2 = iu2 + 1 < sv ? samp[iu2 + 1] : 0xfffffffeu;
        G->d_in = dsamp[il] > dsamp[iu] ? dsamp[il] : dsamp[iu];
        const K d_lo2 = il2 > 0 ? dsamp[il2 - 1] : ~0u, d_up2 = iu2 + 1 < sv ? dsamp[iu2 + 1] : ~0u;
        G->d_out = d_lo2 < d_up2 ? d_lo2 : d_up2;
        G->m = m; G->m2 = m2; G->c = c;
        G->cursor = 0; G->below = 0;
    }
}

// ------------------------------------------------------------------------------------------
// group: MAD from the candidate rings, thresholds, exact raw-domain thresholds
__global__ void __launch_bounds__(kBigNT, 3)
big_mad_kernel(BigGeom g, BigGroup* __restrict__ groups, uint32_t* __restrict__ cand_all,
               rfi_tile_stat_t* __restrict__ stats, int* __restrict__ fail_list, int* __restrict__ fail_count) {
    using T = float;
    using K = uint32_t;
    constexpr int NT = kBigNT;
    constexpr K kInfKey = 0x7f800000u, kExcl = ~0u;
    __shared__ MonoShared<K> sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long grp = blockIdx.x;
    const PlanDev& pp = g.p;
    BigGroup* G = groups + grp;
    if (G->fail) return;
    const uint32_t nv = (uint32_t)g.P * (uint32_t)g.P, k1 = (nv - 1) >> 1, k2 = nv >> 1;
    const uint32_t M = G->cursor, B = G->below;
    const T m = G->m, m2 = G->m2, c = G->c;
    const K d_in = G->d_in, d_out = G->d_out;
    if (M > g.cap || B > k1 || k2 >= B + M) { big_fail(groups, grp, 9, fail_list, fail_count); return; }
    K* cand = cand_all + (size_t)grp * g.cap;
    for (uint32_t i = tid; i < M; i += NT) {
        const T ps = proc_nofill<T>(raw_val<T>(cand[i]), pp, m, m2);
        cand[i] = to_key<T>(fabs_(ps - c));
    }
    __syncthreads();
    K r1k, r2k;
    mono_resolve<K, NT>(cand, M, k1 - B, k2 - B, r1k, r2k, sh);
    if (r1k < d_in || r2k > d_out) { big_fail(groups, grp, 11, fail_list, fail_count); return; }
    const T d = median_of_pair<T>(from_key<T>(r1k), from_key<T>(r2k), nv);
    const T ds = d * (T)pp.sigma;
    const T thr_hi = c + ds, thr_lo = c - ds;
    // exact raw thresholds (same search as the P = 128 kernel): warp 0 raw_hi, warp 1 raw_lo
    if (warp < 2) {
        constexpr K kTop = kInfKey;
        const bool want_hi = (warp == 0);
        const T thr = want_hi ? thr_hi : thr_lo;
        auto pre = [&](K k) {
            const T v = proc_nofill<T>(raw_val<T>(k), pp, m, m2);
            return want_hi ? (v <= thr) : !(v >= thr);
        };
        K first_false;
        if (!(thr == thr)) first_false = want_hi ? kTop : K(0);
        else if (!pre(K(0))) first_false = 0;
        else if (pre(kTop - 1)) first_false = kTop;
        else {
            K a = 0, b = kTop - 1;
            T gg = thr;
            if (pp.norm_after && m2 > T(0)) gg = gg * m2;
            if (pp.stretch == RFI_STRETCH_SQRT) gg = gg * gg;
            else if (pp.stretch == RFI_STRETCH_LOG10) gg = (T)exp10((double)gg);
            if (pp.norm_before && m > T(0)) gg = gg * m;
            K gk = (gg == gg && gg > T(0)) ? __float_as_uint(gg) : K(16);
            gk = gk < K(16) ? K(16) : (gk > kTop - 17 ? kTop - 17 : gk);
            {
                const K tk = gk - 16 + (K)lane;
                const uint32_t bal = __ballot_sync(0xffffffffu, pre(tk));
                if ((bal & 1u) && !(bal >> 31)) {
                    const int nn = __popc(bal);
                    a = gk - 16 + (K)(nn - 1);
                    b = a + 1;
                }
            }
            while (b - a > 1) {
                const K step = (b - a + 32) / 33;
                const K tk = a + (K)(lane + 1) * step;
                const bool good = (tk < b) && pre(tk);
                const int nn = __popc(__ballot_sync(0xffffffffu, good));
                const K na = a + (K)nn * step, nb = a + (K)(nn + 1) * step;
                a = na;
                b = nb < b ? nb : b;
            }
            first_false = b;
        }
        if (lane == 0) {
            if (want_hi) sh.res2 = (first_false == 0) ? kExcl : first_false - 1;
            else sh.res1 = first_false;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const K klo = sh.res1, khi = sh.res2;
        const T raw_lo = raw_val<T>(klo);
        const T raw_hi = (khi == kExcl) ? T(-1) : raw_val<T>(khi);
        big_write_stat(stats + grp, pp, m, m2, c, d, thr_lo, thr_hi, raw_lo, raw_hi, (int)nv);
    }
}

// ------------------------------------------------------------------------------------------
// phase-2 arithmetic shared by big_range and big_write (identical ops -> consistent min / max)
struct BigMath {
    float med, fill, med2, thr_lo, thr_hi, raw_lo, raw_hi;
    FastChain chain;
    bool fast;   // monotone group: raw thresholds, approximate log amplitude (see rfi_tiles.cu phase 2)
};

RFI_DEVINL BigMath big_math(const PlanDev& p, const rfi_tile_stat_t& st) {
    BigMath b;
    b.med = (float)st.median_before; b.fill = (float)st.inf_fill; b.med2 = (float)st.median_after;
    b.thr_lo = (float)st.thr_lo; b.thr_hi = (float)st.thr_hi;
    b.raw_lo = (float)st.raw_lo; b.raw_hi = (float)st.raw_hi;
    b.chain = make_fast_chain(p, b.med, b.med2);
    b.fast = (st.route & RFI_TILE_RAW_THRESHOLDS) != 0;
    return b;
}

// log amplitude of one raw sample and its MAD label
template <bool kFast>
RFI_DEVINL float big_eval(float a, const PlanDev& p, const BigMath& b, unsigned char& f) {
    if constexpr (kFast) {
        f = ((a > b.raw_hi) || (a < b.raw_lo)) ? 1 : 0;
        return fast_log_amp(a, b.chain);
    } else {
        const float x = process_sample<float>(a, p, b.med, b.fill, b.med2);
        f = ((x > b.thr_hi) || (x < b.thr_lo)) ? 1 : 0;
        return log10_img(fabsf(x) + 1e-10f);
    }
}

// bulk asynchronous copy shared -> global (cp.async.bulk, SASS UBLKCP), as in the P = 128 writer (rfi_tiles.cu)
RFI_DEVINL void big_bulk_store(void* dst_global, const void* src_shared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                 :: "l"(dst_global), "r"((uint32_t)__cvta_generic_to_shared(src_shared)), "r"(bytes) : "memory");
}
RFI_DEVINL void big_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
RFI_DEVINL void big_fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kBigLP = kP + 1;   // pitch of the log-amplitude tile: conflict-free rows AND columns
constexpr int kBigFP = kP + 4;   // pitch of the transposed label tile (bytes)

// Pass A of a sub-tile: log-amplitude tile, halo rows / columns from the neighbouring sub-tiles of
// the same group, labels.  kWrite: label rows of rotations 0 / 1 go straight to global memory and
// the transposed label tile is kept for rotations 2 / 3; else only the flagged samples are counted.
// kCplx (complex branch, preprocessor.py:562-606): `src` is the complex64 cube itself, L = log10(|z| + 1e-10) with no
// normalisation or stretch, and the phase channel ((angle + pi) / (2 pi), ImageNet-normalised) goes to the tile `Ph`.
template <bool kCplx> struct BigSrc { using type = float; };
template <> struct BigSrc<true> { using type = float2; };

template <bool kFast, bool kCplx>
RFI_DEVINL float big_sample_L(const typename BigSrc<kCplx>::type& v, const PlanDev& p, const BigMath& bm, unsigned char& fm, float& ph) {
    if constexpr (kCplx) {
        fm = 0;
        const float a = cabs_fast(v.x, v.y);
        const float c2 = (atan2f(v.y, v.x) + 3.141592653589793f) * (float)(1.0 / 6.283185307179586);
        ph = __fmaf_rn(c2, 1.0f / 0.225f, (0.0f - 0.406f) / 0.225f);
        return log10_img(a + 1e-10f);
    } else {
        ph = 0.f;
        return big_eval<kFast>(v, p, bm, fm);
    }
}

template <bool kFast, bool kWrite, bool kCplx = false>
RFI_DEVINL void big_pass_a(const BigGeom& g, const BigTile& t, const BigMath& bm, const void* __restrict__ src_v,
                           const uint8_t* __restrict__ flags, float* Ls, float* halo, unsigned char* FbT,
                           uint8_t* lab0, uint8_t* lab1, uint32_t& nflag, float* Ph = nullptr) {
    using S = typename BigSrc<kCplx>::type;
    const S* __restrict__ src = static_cast<const S*>(src_v);
    constexpr int RS = kBigNT / 32, STEPS = kP / RS, Q = kP / 32;
    const PlanDev& p = g.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t T_ = (size_t)p.times;
    S cur[Q], nxt[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) cur[q] = __ldg(src + t.origin + (size_t)warp * T_ + lane + 32 * q);
    uint32_t nf = 0;
#pragma unroll 1
    for (int s = 0; s < STEPS; ++s) {
        const int row = s * RS + warp;
        if (s + 1 < STEPS) {
#pragma unroll
            for (int q = 0; q < Q; ++q) nxt[q] = __ldg(src + t.origin + (size_t)(row + RS) * T_ + lane + 32 * q);
        }
        unsigned char fl[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            fl[q] = 0;
            if (p.flag_mode == RFI_FLAGS_CUSTOM) fl[q] = __ldg(flags + t.origin + (size_t)row * T_ + lane + 32 * q);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int col = lane + 32 * q;
            unsigned char fm;
            float ph;
            const float L = big_sample_L<kFast, kCplx>(cur[q], p, bm, fm, ph);
            const unsigned char f = (p.flag_mode == RFI_FLAGS_MAD) ? fm : fl[q];
            Ls[row * kBigLP + col] = L;
            if constexpr (kCplx && kWrite) Ph[row * kBigLP + col] = ph;
            if constexpr (kWrite) {
                FbT[col * kBigFP + row] = f;
                if (lab0) lab0[(size_t)row * g.P + col] = f;
                if (lab1) lab1[(size_t)(kP - 1 - row) * g.P + col] = f;
            } else {
                nf += f ? 1u : 0u;
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) cur[q] = nxt[q];
    }
    nflag = nf;
    {   // halos: side 0 top row, 1 bottom row, 2 left column, 3 right column
        const int side = tid >> 7, k = tid & 127;
        const bool has = side == 0 ? t.hasT : side == 1 ? t.hasB : side == 2 ? t.hasL : t.hasR;
        float L = 0.f;
        if (has) {
            const size_t at = side == 0 ? t.origin - T_ + k : side == 1 ? t.origin + (size_t)kP * T_ + k
                            : side == 2 ? t.origin + (size_t)k * T_ - 1 : t.origin + (size_t)k * T_ + kP;
            unsigned char fm;
            float ph;
            L = big_sample_L<kFast, kCplx>(__ldg(src + at), p, bm, fm, ph);
        }
        halo[side * kP + k] = L;
    }
}

// Pass B of a sub-tile: NaN-ignoring min / max of the squared gradient of the three distinct
// rotation variants ([0] backward/backward, [1] forward rows, [2] forward columns) and of L ([3]),
// per thread; differences reach into the halo where the neighbouring sub-tile is in the group.
RFI_DEVINL void big_pass_b(const BigTile& t, int R, const float* Ls, const float* halo,
                           float (&lo4)[4], float (&hi4)[4]) {
    constexpr int RS = kBigNT / 32, STEPS = kP / RS, Q = kP / 32, LP = kBigLP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* hT = halo, *hB = halo + kP, *hL = halo + 2 * kP, *hR = halo + 3 * kP;
#pragma unroll
    for (int k = 0; k < 4; ++k) lo4[k] = hi4[k] = Scalar<float>::nan();
#pragma unroll 1
    for (int s = 0; s < STEPS; ++s) {
        const int i = s * RS + warp;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = lane + 32 * q;
            const float c = Ls[i * LP + j];
            const float bi = (i > 0) ? c - Ls[(i - 1) * LP + j] : (t.hasT ? c - hT[j] : 0.f);
            const float bj = (j > 0) ? c - Ls[i * LP + j - 1] : (t.hasL ? c - hL[i] : 0.f);
            const float bi2 = bi * bi, bj2 = bj * bj;
            const float ss0 = __fmaf_rn(bi, bi, bj2);
            lo4[0] = fminf(lo4[0], ss0); hi4[0] = fmaxf(hi4[0], ss0);
            if (R >= 2) {
                const float fi = (i < kP - 1) ? c - Ls[(i + 1) * LP + j] : (t.hasB ? c - hB[j] : 0.f);
                const float ss1 = __fmaf_rn(fi, fi, bj2);
                lo4[1] = fminf(lo4[1], ss1); hi4[1] = fmaxf(hi4[1], ss1);
            }
            if (R >= 4) {
                const float fj = (j < kP - 1) ? c - Ls[i * LP + j + 1] : (t.hasR ? c - hR[i] : 0.f);
                const float ss3 = __fmaf_rn(fj, fj, bi2);
                lo4[2] = fminf(lo4[2], ss3); hi4[2] = fmaxf(hi4[2], ss3);
            }
            lo4[3] = fminf(lo4[3], c); hi4[3] = fmaxf(hi4[3], c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// sub-tile: flag count + min / max of L and of the squared-gradient variants -> group accumulators
template <bool kCplx>
__global__ void __launch_bounds__(kBigNT, 2)
big_range_kernel(BigGeom g, const void* __restrict__ src, const uint8_t* __restrict__ flags,
                 rfi_tile_stat_t* __restrict__ stats, BigGroup* __restrict__ groups,
                 const int* __restrict__ list, int count_flags) {
    constexpr int LP = kBigLP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ls = reinterpret_cast<float*>(smem_raw);   // [kP][LP]
    float* halo = Ls + (size_t)kP * LP;               // [4][kP]
    float* red = halo + 4 * kP;                       // [NT / 32 * 8]
    const long long blk = list ? (long long)list[blockIdx.x / g.n2] * g.n2 + (blockIdx.x % g.n2) : (long long)blockIdx.x;
    const BigTile t = big_tile(g, blk);
    if (!list && groups[t.grp].fail) return;          // measured later, after the fallback
    const PlanDev& p = g.p;
    const int lane = threadIdx.x & 31;
    const rfi_tile_stat_t st = stats[t.grp];
    const BigMath bm = big_math(p, st);
    uint32_t nf = 0;
    if constexpr (kCplx) big_pass_a<false, false, true>(g, t, bm, src, flags, Ls, halo, nullptr, nullptr, nullptr, nf);
    else if (bm.fast) big_pass_a<true, false>(g, t, bm, src, flags, Ls, halo, nullptr, nullptr, nullptr, nf);
    else big_pass_a<false, false>(g, t, bm, src, flags, Ls, halo, nullptr, nullptr, nullptr, nf);
    __syncthreads();
    float lo4[4], hi4[4];
    big_pass_b(t, p.rotations, Ls, halo, lo4, hi4);
    block_nanminmax4<kBigNT, float>(lo4, hi4, red);
    if (count_flags) nf = __reduce_add_sync(0xffffffffu, nf);
    if (count_flags && lane == 0 && nf) atomicAdd(reinterpret_cast<unsigned int*>(&stats[t.grp].n_flagged), nf);
    if (threadIdx.x == 0) {
        uint32_t* rng = groups[t.grp].rng;
        // rng: [0,1] L, [2,3] variant 0, [4,5] variant 1, [6,7] variant 3
        const int at[4] = {2, 4, 6, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (lo4[k] == lo4[k]) atomicMin(&rng[at[k]], to_key<float>(lo4[k]));
            if (hi4[k] == hi4[k]) atomicMax(&rng[at[k] + 1], to_key<float>(hi4[k]));
        }
    }
}

// sub-tile: flagged samples only (the cluster writer finds the ranges itself): two compares on
// the raw magnitude, or the caller's flag bytes.  Groups measured by the generic select were
// counted there.
__global__ void __launch_bounds__(kBigNT)
big_count_kernel(BigGeom g, const float* __restrict__ src, const uint8_t* __restrict__ flags,
                 rfi_tile_stat_t* __restrict__ stats, const BigGroup* __restrict__ groups) {
    constexpr int G8 = kP * kP / kBigNT / 4, RS = kBigNT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BigTile t = big_tile(g, blockIdx.x);
    if (groups[t.grp].fail) return;
    uint32_t nf = 0;
    if (g.p.flag_mode == RFI_FLAGS_MAD) {
        const float raw_lo = (float)stats[t.grp].raw_lo, raw_hi = (float)stats[t.grp].raw_hi;
#pragma unroll
        for (int g8 = 0; g8 < G8; ++g8) {
            const size_t idx = t.origin + (size_t)(g8 * RS + warp) * g.p.times + lane * 4;
            const float4 q = __ldg(reinterpret_cast<const float4*>(src + idx));
            nf += ((q.x > raw_hi) || (q.x < raw_lo)) ? 1u : 0u;
            nf += ((q.y > raw_hi) || (q.y < raw_lo)) ? 1u : 0u;
            nf += ((q.z > raw_hi) || (q.z < raw_lo)) ? 1u : 0u;
            nf += ((q.w > raw_hi) || (q.w < raw_lo)) ? 1u : 0u;
        }
    } else if (g.p.flag_mode == RFI_FLAGS_CUSTOM) {
#pragma unroll
        for (int g8 = 0; g8 < G8; ++g8) {
            const size_t idx = t.origin + (size_t)(g8 * RS + warp) * g.p.times + lane * 4;
            const uint32_t f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + idx));
            nf += __popc((((f4 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | f4) & 0x80808080u);
        }
    }
    __shared__ uint32_t tot;
    if (tid == 0) tot = 0;
    __syncthreads();
    nf = __reduce_add_sync(0xffffffffu, nf);
    if (lane == 0 && nf) atomicAdd(&tot, nf);
    __syncthreads();
    if (tid == 0 && tot) atomicAdd(reinterpret_cast<unsigned int*>(&stats[t.grp].n_flagged), tot);
}

RFI_DEVINL ChanScale<float> big_scale(uint32_t kmin, uint32_t kmax, bool take_sqrt) {
    float lo = (kmin == ~0u) ? Scalar<float>::nan() : from_key<float>(kmin);
    float hi = (kmax == 0u) ? Scalar<float>::nan() : from_key<float>(kmax);
    if (take_sqrt) { lo = sqrt_fast(lo); hi = sqrt_fast(hi); }
    return make_scale<float>(lo, hi);
}

// ------------------------------------------------------------------------------------------
// phase 2: one CTA per sub-tile, every kept rotation written into its block of the output patch
//
// kCluster: the n2 CTAs of a group form one thread-block cluster; each finds the min / max of its
// own sub-tile after pass A and the group's ranges are combined through distributed shared memory
// (one cluster barrier pair), so no separate range launch reads the magnitudes again.  Otherwise
// the ranges come from big_range_kernel's accumulators.
template <bool kCluster, bool kCplx = false>
__global__ void __launch_bounds__(kBigNT, kCplx ? 1 : 2)
big_write_kernel(BigGeom g, const void* __restrict__ src, const uint8_t* __restrict__ flags,
                 const rfi_tile_stat_t* __restrict__ stats, const BigGroup* __restrict__ groups,
                 const long long* __restrict__ dest_slot, float* __restrict__ images,
                 uint8_t* __restrict__ labels) {
    constexpr int NT = kBigNT, RS = NT / 32, STEPS = kP / RS, Q = kP / 32, LP = kBigLP, FP = kBigFP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ls = reinterpret_cast<float*>(smem_raw);                       // [kP][LP]
    float* halo = Ls + (size_t)kP * LP;                                   // [4][kP]
    unsigned char* FbT = reinterpret_cast<unsigned char*>(halo + 4 * kP); // [kP][FP]
    float* stage = reinterpret_cast<float*>(FbT + (size_t)kP * FP);       // [warps][3 * kP]
    [[maybe_unused]] float* Ph = stage + (size_t)(NT / 32) * 3 * kP;      // [kP][LP] phase channel (complex branch)

    const BigTile t = big_tile(g, blockIdx.x);
    const PlanDev& p = g.p;
    const int per = p.nh * p.nw, R = p.rotations, n = g.n, P = g.P;
    const long long base = t.w * R * per;
    long long slot0 = dest_slot[base + (long long)t.TI * p.nw + t.TJ], slot1 = -1, slot2 = -1, slot3 = -1;
    if (R >= 2) slot1 = dest_slot[base + per + (long long)(p.nh - 1 - t.TI) * p.nw + t.TJ];
    if (R >= 4) {
        slot2 = dest_slot[base + 2LL * per + (long long)t.TJ * p.nh + t.TI];
        slot3 = dest_slot[base + 3LL * per + (long long)(p.nw - 1 - t.TJ) * p.nh + t.TI];
    }
    if (slot0 < 0 && slot1 < 0 && slot2 < 0 && slot3 < 0) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const rfi_tile_stat_t st = stats[t.grp];
    const BigMath bm = big_math(p, st);
    const size_t PP = (size_t)P * P;
    // block of this sub-tile inside each rotated patch: (row block, column block)
    const int br[4] = {t.si, n - 1 - t.si, t.sj, n - 1 - t.sj};
    const int bc[4] = {t.sj, t.sj, t.si, t.si};
    auto lab_at = [&](long long sl, int r) -> uint8_t* {
        return sl < 0 ? nullptr : labels + (size_t)sl * PP + (size_t)br[r] * kP * P + (size_t)bc[r] * kP;
    };
    uint32_t nf_unused = 0;
    if constexpr (kCplx) big_pass_a<false, true, true>(g, t, bm, src, flags, Ls, halo, FbT, lab_at(slot0, 0), lab_at(slot1, 1), nf_unused, Ph);
    else if (bm.fast) big_pass_a<true, true>(g, t, bm, src, flags, Ls, halo, FbT, lab_at(slot0, 0), lab_at(slot1, 1), nf_unused);
    else big_pass_a<false, true>(g, t, bm, src, flags, Ls, halo, FbT, lab_at(slot0, 0), lab_at(slot1, 1), nf_unused);
    __syncthreads();

    ChanScale<float> ls, g0, g1, g3;
    if constexpr (kCluster) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        __shared__ float xch[8];   // this CTA's (lo, hi) x 4, read by every CTA of the cluster
        float lo4[4], hi4[4];
        big_pass_b(t, R, Ls, halo, lo4, hi4);
        block_nanminmax4<NT, float>(lo4, hi4, stage);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { xch[k] = lo4[k]; xch[4 + k] = hi4[k]; }
        }
        cluster.sync();
        const unsigned nc = cluster.num_blocks();
        for (unsigned r = 0; r < nc; ++r) {
            if (r == cluster.block_rank()) continue;
            const float* peer = cluster.map_shared_rank(xch, r);
#pragma unroll
            for (int k = 0; k < 4; ++k) { lo4[k] = fminf(lo4[k], peer[k]); hi4[k] = fmaxf(hi4[k], peer[4 + k]); }
        }
        cluster.sync();            // no CTA leaves (or reuses xch) while a peer still reads it
        g0 = make_scale<float>(sqrt_fast(lo4[0]), sqrt_fast(hi4[0]));
        g1 = make_scale<float>(sqrt_fast(lo4[1]), sqrt_fast(hi4[1]));
        g3 = make_scale<float>(sqrt_fast(lo4[2]), sqrt_fast(hi4[2]));
        ls = make_scale<float>(lo4[3], hi4[3]);
    } else {
        const uint32_t* rng = groups[t.grp].rng;
        ls = big_scale(rng[0], rng[1], false);
        g0 = big_scale(rng[2], rng[3], true);
        g1 = big_scale(rng[4], rng[5], true);
        g3 = big_scale(rng[6], rng[7], true);
    }

    const float mean0 = 0.485f, mean1 = 0.456f, mean2 = 0.406f;
    const float std0 = 0.229f, std1 = 0.224f, std2 = 0.225f;
    const float is0 = 1.0f / std0, is1 = 1.0f / std1;
    const float nb0 = (0.0f - mean0) / std0, nb1 = (0.0f - mean1) / std1, nb2 = (0.0f - mean2) / std2;
    const float* hT = halo, *hB = halo + kP, *hL = halo + 2 * kP, *hR = halo + 3 * kP;
    float* wstage = stage + (size_t)warp * 3 * kP;

    // kPlain / the constant third channel: as in the P = 128 writer (rfi_tiles.cu, write_tile)
    if constexpr (!kCplx) {
#pragma unroll
        for (int q = 0; q < Q; ++q) wstage[(lane + 32 * q) * 3 + 2] = nb2;
        __syncwarp();
    }
    auto emit_as = [&](auto rot_tag, auto plain_tag, long long sl, const ChanScale<float>& gs) {
        constexpr int rot = decltype(rot_tag)::value;
        constexpr bool kPlain = decltype(plain_tag)::value;
        float* out_img = images + ((size_t)sl * PP + (size_t)br[rot] * kP * P + (size_t)bc[rot] * kP) * 3;
        [[maybe_unused]] unsigned char* out_lab = labels + (size_t)sl * PP + (size_t)br[rot] * kP * P + (size_t)bc[rot] * kP;
        constexpr int srow = (rot == 0) ? LP : (rot == 1) ? -LP : (rot == 2) ? 1 : -1;
        constexpr int scol = (rot <= 1) ? 1 : LP;
        constexpr int base_at = (rot == 0) ? 0 : (rot == 1) ? (kP - 1) * LP : (rot == 2) ? 0 : (kP - 1);
        constexpr int db = -scol;
        // halo behind output row 0 (indexed by the output column) and behind output column 0
        // (indexed by the output row, reversed for the flipped rotations)
        const float* hprev = (rot == 0) ? hT : (rot == 1) ? hB : (rot == 2) ? hL : hR;
        const bool has_prev = (rot == 0) ? t.hasT : (rot == 1) ? t.hasB : (rot == 2) ? t.hasL : t.hasR;
        const float* hleft = (rot <= 1) ? hL : hT;
        const bool has_left = (rot <= 1) ? t.hasL : t.hasT;
        const int row0 = warp * STEPS;
        const float ga = gs.inv * is0, gb = __fmaf_rn(-gs.lo * gs.inv, is0, nb0);
        const float la = ls.inv * is1, lb = __fmaf_rn(-ls.lo * ls.inv, is1, nb1);
        float prev[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int ocol = lane + 32 * q;
            const int at = base_at + (row0 - 1) * srow + ocol * scol;
            prev[q] = (row0 > 0) ? Ls[at] : hprev[ocol];
        }
#pragma unroll 1
        for (int s = 0; s < STEPS; ++s) {
            const int orow = row0 + s;
            const float left0 = hleft[(rot == 1 || rot == 3) ? (kP - 1 - orow) : orow];
            float o[Q][3];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int ocol = lane + 32 * q;
                const int at = base_at + orow * srow + ocol * scol;
                const float c = Ls[at];
                const float td = (orow > 0 || has_prev) ? c - prev[q] : 0.f;
                const float fd = (ocol > 0) ? c - Ls[at + db] : (has_left ? c - left0 : 0.f);
                prev[q] = c;
                const float gr = sqrt_fast(__fmaf_rn(td, td, fd * fd));
                o[q][0] = (kPlain || gs.ok) ? __fmaf_rn(gr, ga, gb) : nb0;   // flat / all-NaN channel: exactly 0 before ImageNet
                if constexpr (kCplx) {   // fixed scale clip((L + 3) / 7, 0, 1) and the phase (preprocessor.py:588-604)
                    float u = (c - (-3.0f)) * (float)(1.0 / 7.0);
                    u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);  // np.clip keeps NaN
                    o[q][1] = __fmaf_rn(u, is1, nb1);
                    o[q][2] = Ph[at];
                } else {
                    o[q][1] = (kPlain || ls.ok) ? __fmaf_rn(c, la, lb) : nb1;
                    o[q][2] = nb2;
                }
            }
            // the staged row (1536 contiguous bytes of the output row) leaves as one bulk asynchronous copy;
            // the previous row's copy has read the staging buffer by the time this row's values exist
            if (lane == 0) big_bulk_wait_read();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int ocol = lane + 32 * q;
                wstage[ocol * 3 + 0] = o[q][0];
                wstage[ocol * 3 + 1] = o[q][1];
                if constexpr (kCplx) wstage[ocol * 3 + 2] = o[q][2];
            }
            big_fence_async_shared();
            __syncwarp();
            if (lane == 0) big_bulk_store(out_img + (size_t)orow * P * 3, wstage, 3 * kP * sizeof(float));
            if constexpr (rot >= 2) {
                const unsigned char* lrow = FbT + ((rot == 2) ? orow : (kP - 1 - orow)) * FP;
                reinterpret_cast<uint32_t*>(out_lab + (size_t)orow * P)[lane] = reinterpret_cast<const uint32_t*>(lrow)[lane];
            }
        }
    };
    auto emit = [&](auto rot_tag, long long sl, const ChanScale<float>& gs) {
        if (sl < 0) return;  // uniform across the block
        if (gs.ok && (kCplx || ls.ok)) emit_as(rot_tag, std::true_type{}, sl, gs);
        else emit_as(rot_tag, std::false_type{}, sl, gs);
    };
    emit(std::integral_constant<int, 0>{}, slot0, g0);
    emit(std::integral_constant<int, 1>{}, slot1, g1);
    emit(std::integral_constant<int, 2>{}, slot2, g0);
    emit(std::integral_constant<int, 3>{}, slot3, g3);
    if (lane == 0) big_bulk_wait_read();  // shared memory must outlive the last copy's read
}

// ------------------------------------------------------------------------------------------
// host side
bool plan_is_big(const rfi_plan_t* plan) {
    if (!plan) return false;
    const int P = plan->patch;
    if (P != 256 && P != 512 && P != 1024) return false;
    if (plan->channels % P || plan->times % P || plan->channels < P || plan->times < P) return false;
    // (a waterfall of exactly P x P is not patchified by the reference, preprocessor.py:261 -- the "patches" are
    //  then the R rotated waterfalls, which is what one tile per waterfall gives)
    if (plan->dtype != RFI_F32 && plan->dtype != RFI_C64) return false;
    // complex branch (no normalisation, no stretch): with the caller's flags or none; MAD flags of |z| stay generic
    if (plan->dtype == RFI_C64 && !plan->magnitude && plan->flag_mode == RFI_FLAGS_MAD) return false;
    const bool need_median = plan->norm_before || plan->norm_after || plan->flag_mode == RFI_FLAGS_MAD;
    if (!need_median && plan->stretch != RFI_STRETCH_NONE) return false;  // inf fill without a median pass
    if (plan->rotations != 1 && plan->rotations != 2 && plan->rotations != 4) return false;
    if (plan->stretch < 0 || plan->stretch > 2 || plan->flag_mode < 0 || plan->flag_mode > 2) return false;
    if (plan->n_waterfalls < 0) return false;
    const long long groups = plan->n_waterfalls * (plan->channels / P) * (plan->times / P);
    if (groups * (long long)((P / kP) * (P / kP)) > 0x7fffffffLL) return false;
    return true;
}

static void make_big(const rfi_plan_t* plan, BigGeom& g) {
    PlanDev& d = g.p;
    const int P = plan->patch;
    d.n_waterfalls = plan->n_waterfalls; d.channels = plan->channels; d.times = plan->times;
    d.nh = (int)(plan->channels / P); d.nw = (int)(plan->times / P);
    d.rotations = plan->rotations; d.stretch = plan->stretch;
    d.norm_before = plan->norm_before; d.norm_after = plan->norm_after;
    d.flag_mode = plan->flag_mode; d.magnitude = plan->magnitude; d.sigma = plan->sigma;
    g.P = P; g.n = P / kP; g.n2 = g.n * g.n;
    g.n_groups = plan->n_waterfalls * d.nh * d.nw;
    g.cap = (uint32_t)((long long)P * P / 4);
    g.delta = (int)(2.25f * sqrtf((float)kBigS)) + 2;  // 4.5 sigma of a sample rank
}

struct BigWs {
    void* generic;
    BigGroup* groups;
    int* fail_count;
    int* fail_list;
    uint32_t* samples;
    uint32_t* cand;
    float* mag;
    size_t bytes;
};

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

static BigWs big_ws(const rfi_plan_t* plan, const BigGeom& g, void* ws) {
    BigWs w;
    char* b = static_cast<char*>(ws);
    size_t off = 0;
    w.generic = b + off; off += up256(generic_workspace_bytes(plan));
    w.groups = reinterpret_cast<BigGroup*>(b + off); off += up256((size_t)g.n_groups * sizeof(BigGroup));
    w.fail_count = reinterpret_cast<int*>(b + off); off += 256;
    w.fail_list = reinterpret_cast<int*>(b + off); off += up256((size_t)g.n_groups * sizeof(int));
    w.samples = reinterpret_cast<uint32_t*>(b + off); off += up256((size_t)g.n_groups * kBigS * sizeof(uint32_t));
    w.cand = reinterpret_cast<uint32_t*>(b + off); off += up256((size_t)g.n_groups * g.cap * sizeof(uint32_t));
    w.mag = reinterpret_cast<float*>(b + off);
    if (plan->dtype == RFI_C64 && plan->magnitude) off += up256((size_t)plan->n_waterfalls * plan->channels * plan->times * sizeof(float));
    w.bytes = off;
    return w;
}

size_t big_workspace_bytes(const rfi_plan_t* plan) {
    BigGeom g;
    make_big(plan, g);
    return big_ws(plan, g, nullptr).bytes;
}

// P = 256: the four sub-tile CTAs of a group run as one cluster (RFI_BIG_NO_CLUSTER=1 keeps the
// range launch instead -- for A/B timing)
static bool big_use_cluster(const BigGeom& g) {
    static const bool off = getenv("RFI_BIG_NO_CLUSTER") != nullptr;
    return g.n2 == 4 && !off;
}

static size_t big_range_smem() { return ((size_t)kP * kBigLP + 4 * kP + kBigNT / 32 * 8) * sizeof(float); }
static size_t big_write_smem(bool cplx = false) {
    return ((size_t)kP * kBigLP + 4 * kP) * sizeof(float) + (size_t)kP * kBigFP + (size_t)(kBigNT / 32) * 3 * kP * sizeof(float) +
           (cplx ? (size_t)kP * kBigLP * sizeof(float) : 0);
}

// complex branch: nothing is measured on the data; the statistics entry only carries the flag count
__global__ void big_neutral_stats_kernel(BigGeom g, rfi_tile_stat_t* __restrict__ stats) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n_groups) return;
    rfi_tile_stat_t st;
    st.median_before = st.inf_fill = st.median_after = 0.0;
    st.centre = st.mad = st.thr_lo = st.thr_hi = 0.0;
    st.n_valid = g.P * g.P; st.n_inf = 0; st.n_flagged = 0; st.route = 0; st.raw_lo = st.raw_hi = 0.0;
    stats[i] = st;
}

int big_tile_stats(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                   rfi_tile_stat_t* stats, void* workspace, cudaStream_t st) {
    BigGeom g;
    make_big(plan, g);
    if (g.n_groups == 0) return RFI_OK;
    if (!data || !stats) { set_error("data / stats is NULL"); return RFI_E_INVALID; }
    if (!workspace) { set_error("patch sizes above 128 need the workspace rfi_plan_workspace_bytes() reports"); return RFI_E_INVALID; }
    if (g.p.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    const BigWs w = big_ws(plan, g, workspace);
    if (plan->dtype == RFI_C64 && !plan->magnitude) {
        // complex branch (preprocessor.py:285-292: no normalisation, no stretch): flag counts per patch, and --
        // unless the writer's cluster finds them itself -- the per-patch ranges of the gradient variants
        const unsigned subs_c = (unsigned)(g.n_groups * g.n2), groups_c = (unsigned)g.n_groups;
        const int count = g.p.flag_mode == RFI_FLAGS_CUSTOM ? 1 : 0;
        big_init_kernel<<<(groups_c + 255) / 256, 256, 0, st>>>(g, w.groups, w.fail_count);
        big_neutral_stats_kernel<<<(groups_c + 255) / 256, 256, 0, st>>>(g, stats);
        if (big_use_cluster(g)) {
            if (count) big_count_kernel<<<subs_c, kBigNT, 0, st>>>(g, nullptr, flags, stats, w.groups);
        } else {
            RFI_CUDA_TRY(cudaFuncSetAttribute(big_range_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_range_smem()));
            big_range_kernel<true><<<subs_c, kBigNT, big_range_smem(), st>>>(g, data, flags, stats, w.groups, nullptr, count);
        }
        RFI_CUDA_TRY(cudaGetLastError());
        return RFI_OK;
    }
    const bool cplx = plan->dtype == RFI_C64;
    const bool need_median = g.p.norm_before || g.p.norm_after || g.p.flag_mode == RFI_FLAGS_MAD;
    const float* src = cplx ? w.mag : static_cast<const float*>(data);
    const unsigned subs = (unsigned)(g.n_groups * g.n2), groups = (unsigned)g.n_groups;
    const bool cluster = big_use_cluster(g);

    big_init_kernel<<<(groups + 255) / 256, 256, 0, st>>>(g, w.groups, w.fail_count);
    if (cplx) big_load_kernel<RFI_C64><<<subs, kBigNT, 0, st>>>(g, data, w.mag, w.groups, w.samples, need_median ? 1 : 0);
    else if (need_median) big_load_kernel<RFI_F32><<<subs, kBigNT, 0, st>>>(g, data, w.mag, w.groups, w.samples, 1);
    int n_fail = 0;
    if (need_median) {
        big_sample_kernel<<<groups, 128, 0, st>>>(g, w.groups, w.samples, w.fail_list, w.fail_count);
        big_pass_kernel<0><<<subs, kBigNT, 0, st>>>(g, src, w.groups, w.cand);
        big_median_kernel<<<groups, kBigNT, 0, st>>>(g, w.groups, w.samples, w.cand, stats, w.fail_list, w.fail_count);
        if (g.p.flag_mode == RFI_FLAGS_MAD) {
            big_pass_kernel<1><<<subs, kBigNT, 0, st>>>(g, src, w.groups, w.cand);
            big_mad_kernel<<<groups, kBigNT, 0, st>>>(g, w.groups, w.cand, stats, w.fail_list, w.fail_count);
        }
        RFI_CUDA_TRY(cudaGetLastError());
        // groups the sampled brackets could not settle (or with NaN / inf / negative / inf-filled
        // samples) are measured by the generic select; the count decides whether it is launched
        RFI_CUDA_TRY(cudaMemcpyAsync(&n_fail, w.fail_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        RFI_CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        // no statistic needed (custom flags / inference without normalisation or stretch): every
        // group takes the exact log amplitude; the generic init writes the neutral statistics
        // through the same subset entry point with the identity list
        n_fail = -1;
    }
    RFI_CUDA_TRY(cudaFuncSetAttribute(big_range_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_range_smem()));
    if (n_fail != 0) {
        if (n_fail < 0) {
            // identity list
            std::vector<int> ident((size_t)g.n_groups);
            for (size_t i = 0; i < ident.size(); ++i) ident[i] = (int)i;
            RFI_CUDA_TRY(cudaMemcpyAsync(w.fail_list, ident.data(), ident.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            RFI_CUDA_TRY(cudaStreamSynchronize(st));
            n_fail = (int)g.n_groups;
        }
        int rc = generic_tile_stats_subset(plan, data, flags, stats, w.generic, w.fail_list, n_fail, st);
        if (rc) return rc;
        if (!cluster) {
            big_rearm_kernel<<<(n_fail + 255) / 256, 256, 0, st>>>(w.groups, w.fail_list, n_fail);
            // ranges of the listed groups (their flag counts come from the generic path)
            big_range_kernel<false><<<(unsigned)n_fail * g.n2, kBigNT, big_range_smem(), st>>>(g, src, flags, stats, w.groups, w.fail_list, 0);
        }
    }
    if (need_median) {
        const int count = g.p.flag_mode != RFI_FLAGS_INFERENCE ? 1 : 0;
        if (!cluster) big_range_kernel<false><<<subs, kBigNT, big_range_smem(), st>>>(g, src, flags, stats, w.groups, nullptr, count);
        else if (count) big_count_kernel<<<subs, kBigNT, 0, st>>>(g, src, flags, stats, w.groups);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

int big_write_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                      const rfi_tile_stat_t* stats, const long long* dest_slot, float* images,
                      uint8_t* labels, void* workspace, cudaStream_t st) {
    BigGeom g;
    make_big(plan, g);
    if (g.n_groups == 0) return RFI_OK;
    if (!data || !stats || !dest_slot) { set_error("data / stats / dest_slot is NULL"); return RFI_E_INVALID; }
    if (!workspace) { set_error("patch sizes above 128 need the workspace rfi_tile_stats was given"); return RFI_E_INVALID; }
    if (g.p.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    const BigWs w = big_ws(plan, g, workspace);
    const bool cb = plan->dtype == RFI_C64 && !plan->magnitude;
    const void* src = cb ? data : (plan->dtype == RFI_C64 ? static_cast<const void*>(w.mag) : data);
    const unsigned subs = (unsigned)(g.n_groups * g.n2);
    if (big_use_cluster(g)) {
        auto kern = cb ? big_write_kernel<true, true> : big_write_kernel<true, false>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_write_smem(cb)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(subs); cfg.blockDim = dim3(kBigNT); cfg.dynamicSmemBytes = big_write_smem(cb); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)g.n2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const BigGroup* groups = w.groups;
        RFI_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, g, src, flags, stats, groups, dest_slot, images, labels));
    } else {
        auto kern = cb ? big_write_kernel<false, true> : big_write_kernel<false, false>;
        RFI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_write_smem(cb)));
        kern<<<subs, kBigNT, big_write_smem(cb), st>>>(g, src, flags, stats, w.groups, dest_slot, images, labels);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

}  // namespace rfi
