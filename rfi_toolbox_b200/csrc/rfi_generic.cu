// rfi_generic.cu -- create_dataset for every geometry the fast path does not take:
// any patch size (256, 512, 1024 ... or odd sizes), waterfalls whose dims are not multiples
// of the patch size (the reference zero-pads bottom/right AFTER the rotation,
// preprocessor.py:527-550) and waterfalls no larger than the patch (patchify skipped,
// preprocessor.py:261-269).
//
// A patch no longer fits one CTA's registers / shared memory here (P = 1024 is 1 Mi samples),
// so the statistics are found by a segmented, multi-pass radix select over global memory:
//
//   group   = the set of samples one median / MAD is taken over.  Without padding the R
//             rotated patches of an original tile share one group (statistics are rotation
//             invariant); with padding every output patch is its own group, because the pad
//             is applied after the flip and each rotation sees a different window.
//   pass    = every CTA re-derives its samples from the cube (magnitude, normalise, stretch --
//             a pure function of the raw sample and the statistics found so far), compares
//             them with 15 trial keys held in registers (4 bits of the answer per pass, no
//             shared-memory histogram atomics) and adds its counts to the group's counters.
//   decide  = one thread per group fixes the next 4 bits.
//
// 8 passes per float32 order statistic, 16 per float64.  All launches are stream ordered;
// groups that do not need a stage (no +-inf sample -> no inf fill) exit at once.
//
// Phase 2 recomputes every output pixel from the cube (three log-amplitude evaluations for
// the backward differences), after a per-patch min/max pass.  Everything numeric goes through
// the same helpers as the fast path (rfi_tiles.cuh), so labels are bit-identical between the
// two paths and to the reference.
#include "rfi_tiles.cuh"

namespace rfi {

constexpr int kGT = 256;         // threads per CTA
constexpr int kGE = 16;          // samples per thread per CTA
constexpr int kGChunk = kGT * kGE;

struct GGeom {
    long long C, T;          // waterfall rows (channels) / cols (times)
    long long n_waterfalls;
    long long n_groups, n_patches;
    int P;                   // patch pitch
    int Pr, Pc;              // window rows / cols in SOURCE orientation (P, P; or C, T when patchify is skipped)
    int nhc, nwc, per;       // blocks along C / along T, per = nhc * nwc
    int R, padded, skip;
    int chunks;              // CTAs per group / per patch
    int stretch, norm_before, norm_after, flag_mode, real_branch;
    double sigma;
    // statistics over a SUBSET of the groups (the big-tile path's fallback, rfi_bigtile.cu):
    // `list` holds n_active group indices; NULL = all n_groups groups
    const int* list;
    long long n_active;
};

RFI_DEVINL long long active_group(const GGeom& g, long long i) { return g.list ? (long long)g.list[i] : i; }

// select state of one group (workspace)
struct GSel {
    unsigned long long prefix, nxt;
    unsigned int cnt[16];    // [0..14] keys below trial t+1; [15] valid keys
    unsigned int cle, n, k1, k2;
    int shift, pad;
    double cf;               // centre of the finite stretched samples (inf fill select)
    unsigned int n_fin, pad2;
};

// per output patch: min / max of the log amplitude and of the gradient magnitude, as ordered keys
struct GRange {
    unsigned long long lmin, lmax, gmin, gmax;
};

enum GStage {
    GS_RAW = 0,       // raw sample                               -> median_before
    GS_FIN = 1,       // finite stretched samples                  -> cf
    GS_FIN_DEV = 2,   // |s - cf| over the finite samples          -> inf_fill
    GS_AFTER = 3,     // stretched, inf filled                     -> median_after
    GS_PROC = 4,      // processed sample                          -> centre
    GS_PROC_DEV = 5,  // |x - centre|                              -> mad, thresholds
};

// window of group g in source coordinates (may stick out of the waterfall: zero pad)
struct GWindow {
    long long w;
    long long r0, c0;
};

RFI_DEVINL GWindow group_window(const GGeom& g, long long grp) {
    GWindow win;
    if (!g.padded) {
        win.w = grp / g.per;
        const int t = (int)(grp % g.per);
        win.r0 = (long long)(t / g.nwc) * g.P;
        win.c0 = (long long)(t % g.nwc) * g.P;
        return win;
    }
    const long long rp = (long long)g.R * g.per;
    win.w = grp / rp;
    const int rem = (int)(grp % rp);
    const int r = rem / g.per, t = rem % g.per;
    if (r <= 1) {
        const int bi = t / g.nwc, bj = t % g.nwc;
        win.r0 = (r == 0) ? (long long)bi * g.P : g.C - (long long)(bi + 1) * g.P;
        win.c0 = (long long)bj * g.P;
    } else {
        const int bi = t / g.nhc, bj = t % g.nhc;  // rotated grid: rows step over T, cols over C
        win.r0 = (long long)bj * g.P;
        win.c0 = (r == 2) ? (long long)bi * g.P : g.T - (long long)(bi + 1) * g.P;
    }
    return win;
}

// output patch q (canonical index) -> its group, rotation and the affine map from patch
// pixel (y, x) to source (row, col):  row = r0 + y * ry + x * rx,  col = c0 + y * cy + x * cx
struct GPatch {
    long long w, grp;
    long long r0, c0;
    int ry, rx, cy, cx;
    int rows, cols;  // patch shape
};

RFI_DEVINL GPatch patch_map(const GGeom& g, long long q) {
    GPatch p;
    const long long rp = (long long)g.R * g.per;
    p.w = q / rp;
    const int rem = (int)(q % rp);
    const int r = rem / g.per, t = rem % g.per;
    const long long rowsP = g.skip ? g.C : g.P, colsP = g.skip ? g.T : g.P;  // window extent along C / T
    p.rows = (r <= 1) ? (int)rowsP : (int)colsP;
    p.cols = (r <= 1) ? (int)colsP : (int)rowsP;
    int bi, bj, ti, tj;  // rotated-grid tile, original-grid tile (unpadded case)
    if (r <= 1) {
        bi = t / g.nwc; bj = t % g.nwc;
        ti = (r == 0) ? bi : g.nhc - 1 - bi; tj = bj;
    } else {
        bi = t / g.nhc; bj = t % g.nhc;
        ti = bj; tj = (r == 2) ? bi : g.nwc - 1 - bi;
    }
    p.grp = g.padded ? q : p.w * g.per + (long long)ti * g.nwc + tj;
    switch (r) {
        case 0:  // X
            p.r0 = (long long)bi * g.P; p.c0 = (long long)bj * g.P; p.ry = 1; p.rx = 0; p.cy = 0; p.cx = 1; break;
        case 1:  // X[::-1, :]
            p.r0 = g.C - 1 - (long long)bi * g.P; p.c0 = (long long)bj * g.P; p.ry = -1; p.rx = 0; p.cy = 0; p.cx = 1; break;
        case 2:  // X.T
            p.r0 = (long long)bj * g.P; p.c0 = (long long)bi * g.P; p.ry = 0; p.rx = 1; p.cy = 1; p.cx = 0; break;
        default:  // X.T[::-1, :]
            p.r0 = (long long)bj * g.P; p.c0 = g.T - 1 - (long long)bi * g.P; p.ry = 0; p.rx = 1; p.cy = -1; p.cx = 0; break;
    }
    return p;
}

// sample (row, col) of waterfall w; zero outside the waterfall (np.pad constant 0)
template <int DT, bool kPhase>
RFI_DEVINL void load_src(const void* data, const GGeom& g, long long w, long long row, long long col,
                         typename In<DT>::T& mag, typename In<DT>::T& ph) {
    using T = typename In<DT>::T;
    if (row < 0 || row >= g.C || col < 0 || col >= g.T) {
        mag = T(0); ph = T(0);  // |0| = 0, angle(0) = 0
        return;
    }
    load1<DT, kPhase>(data, (size_t)((w * g.C + row) * g.T + col), mag, ph);
}

template <typename T>
struct GCtx {   // statistics found so far, in T
    T m, fill, m2, cf, centre, thr_lo, thr_hi;
    bool divide;
};

template <typename T>
RFI_DEVINL GCtx<T> load_ctx(const GGeom& g, const rfi_tile_stat_t& st, const GSel& s) {
    GCtx<T> c;
    c.m = (T)st.median_before; c.fill = (T)st.inf_fill; c.m2 = (T)st.median_after;
    c.cf = (T)s.cf; c.centre = (T)st.centre; c.thr_lo = (T)st.thr_lo; c.thr_hi = (T)st.thr_hi;
    c.divide = g.norm_before && c.m > T(0);
    return c;
}

template <typename T>
RFI_DEVINL T processed(T a, const GGeom& g, const GCtx<T>& c) {
    if (!g.real_branch) return a;
    if (c.divide) a = a / c.m;
    if (g.stretch != RFI_STRETCH_NONE) {
        a = apply_stretch<T>(a, g.stretch);
        if (is_inf(a)) a = c.fill;
    }
    if (g.norm_after && c.m2 > T(0)) a = a / c.m2;
    return a;
}

template <typename T, int STAGE>
RFI_DEVINL typename Scalar<T>::key_t stage_key(T a, const GGeom& g, const GCtx<T>& c) {
    using K = typename Scalar<T>::key_t;
    if constexpr (STAGE == GS_RAW) {
        return to_key<T>(a);
    } else if constexpr (STAGE == GS_FIN || STAGE == GS_FIN_DEV || STAGE == GS_AFTER) {
        T s = a;
        if (c.divide) s = s / c.m;
        s = apply_stretch<T>(s, g.stretch);
        if constexpr (STAGE == GS_AFTER) {
            if (g.stretch != RFI_STRETCH_NONE && is_inf(s)) s = c.fill;
            return to_key<T>(s);
        } else {
            if (is_inf(s)) return ~K(0);
            if constexpr (STAGE == GS_FIN) return to_key<T>(s);
            return to_key<T>(fabs_(s - c.cf));
        }
    } else {
        const T x = processed<T>(a, g, c);
        if constexpr (STAGE == GS_PROC) return to_key<T>(x);
        return to_key<T>(fabs_(x - c.centre));
    }
}

RFI_DEVINL bool stage_active(int stage, const rfi_tile_stat_t& st, const GSel& s) {
    if (stage == GS_FIN || stage == GS_FIN_DEV) return st.n_inf > 0 && s.n_fin > 0;
    return true;
}

RFI_DEVINL unsigned int warp_sum_u32(unsigned int v) { return __reduce_add_sync(0xffffffffu, v); }

// ------------------------------------------------------------------------------------------
// select passes
template <int DT, int STAGE>
__global__ void __launch_bounds__(kGT)
gsel_count_kernel(GGeom g, const void* __restrict__ data, const rfi_tile_stat_t* __restrict__ stats,
                  GSel* __restrict__ sel) {
    using T = typename In<DT>::T;
    using K = typename Scalar<T>::key_t;
    const long long grp = active_group(g, blockIdx.x / g.chunks);
    const int chunk = blockIdx.x % g.chunks;
    const rfi_tile_stat_t st = stats[grp];
    const GSel s = sel[grp];
    if (!stage_active(STAGE, st, s)) return;
    const GCtx<T> ctx = load_ctx<T>(g, st, s);
    const GWindow win = group_window(g, grp);
    const K prefix = (K)s.prefix;
    const int shift = s.shift;
    const uint32_t n = (uint32_t)g.Pr * (uint32_t)g.Pc;
    unsigned int c[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) c[t] = 0;
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;  // < 2^31 (make_geom)
        if (e >= n) break;
        const uint32_t yy = e / (uint32_t)g.Pc, xx = e - yy * (uint32_t)g.Pc;
        T a, ph;
        load_src<DT, false>(data, g, win.w, win.r0 + yy, win.c0 + xx, a, ph);
        const K key = stage_key<T, STAGE>(a, g, ctx);
#pragma unroll
        for (int t = 0; t < 15; ++t) c[t] += (key < (prefix | ((K)(t + 1) << shift))) ? 1u : 0u;
        c[15] += (key != ~K(0)) ? 1u : 0u;
    }
    __shared__ unsigned int sh[16];
    if (threadIdx.x < 16) sh[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const unsigned int v = warp_sum_u32(c[t]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[t], v);
    }
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&sel[grp].cnt[threadIdx.x], sh[threadIdx.x]);
}

__global__ void gsel_begin_kernel(GGeom g, GSel* __restrict__ sel, int key_bits) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= g.n_active) return;
    const long long grp = active_group(g, gi);
    GSel& s = sel[grp];
    s.prefix = 0; s.nxt = ~0ull; s.cle = 0; s.n = 0; s.k1 = s.k2 = 0;
    s.shift = key_bits - 4;
    for (int t = 0; t < 16; ++t) s.cnt[t] = 0;
}

__global__ void gsel_decide_kernel(GGeom g, GSel* __restrict__ sel, int first) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= g.n_active) return;
    const long long grp = active_group(g, gi);
    GSel& s = sel[grp];
    if (first) {
        s.n = s.cnt[15];
        s.k1 = s.n ? (s.n - 1) >> 1 : 0;
        s.k2 = s.n >> 1;
    }
    int d = 0;
    for (int t = 0; t < 15; ++t) d += (s.cnt[t] <= s.k1) ? 1 : 0;
    s.prefix |= (unsigned long long)d << s.shift;
    s.shift -= 4;
    for (int t = 0; t < 16; ++t) s.cnt[t] = 0;
}

template <int DT, int STAGE>
__global__ void __launch_bounds__(kGT)
gsel_next_kernel(GGeom g, const void* __restrict__ data, const rfi_tile_stat_t* __restrict__ stats,
                 GSel* __restrict__ sel) {
    using T = typename In<DT>::T;
    using K = typename Scalar<T>::key_t;
    const long long grp = active_group(g, blockIdx.x / g.chunks);
    const int chunk = blockIdx.x % g.chunks;
    const rfi_tile_stat_t st = stats[grp];
    const GSel s = sel[grp];
    if (!stage_active(STAGE, st, s) || s.k1 == s.k2) return;
    const GCtx<T> ctx = load_ctx<T>(g, st, s);
    const GWindow win = group_window(g, grp);
    const K prefix = (K)s.prefix;
    const long long n = (long long)g.Pr * g.Pc;
    unsigned int cle = 0;
    K nxt = ~K(0);
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;  // < 2^31 (make_geom)
        if (e >= n) break;
        const uint32_t yy = e / (uint32_t)g.Pc, xx = e - yy * (uint32_t)g.Pc;
        T a, ph;
        load_src<DT, false>(data, g, win.w, win.r0 + yy, win.c0 + xx, a, ph);
        const K key = stage_key<T, STAGE>(a, g, ctx);
        cle += (key <= prefix) ? 1u : 0u;
        const K y = key > prefix ? key : ~K(0);
        nxt = y < nxt ? y : nxt;
    }
    cle = warp_sum_u32(cle);
    nxt = warp_min(nxt);
    if ((threadIdx.x & 31) == 0) {
        if (cle) atomicAdd(&sel[grp].cle, cle);
        atomicMin(&sel[grp].nxt, (unsigned long long)nxt);
    }
}

template <typename T, int STAGE>
__global__ void gsel_finish_kernel(GGeom g, rfi_tile_stat_t* __restrict__ stats, GSel* __restrict__ sel) {
    using K = typename Scalar<T>::key_t;
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= g.n_active) return;
    const long long grp = active_group(g, gi);
    rfi_tile_stat_t& st = stats[grp];
    GSel& s = sel[grp];
    if (!stage_active(STAGE, st, s)) return;
    T med = Scalar<T>::nan();
    if (s.n > 0) {
        const T a = from_key<T>((K)s.prefix);
        T b = a;
        if (s.k2 != s.k1) b = (s.k2 < s.cle) ? a : from_key<T>((K)s.nxt);
        med = median_of_pair<T>(a, b, s.n);
    }
    if (STAGE == GS_RAW) { st.median_before = (double)med; st.n_valid = (int)s.n; }
    else if (STAGE == GS_FIN) s.cf = (double)med;
    else if (STAGE == GS_FIN_DEV) st.inf_fill = (double)med;
    else if (STAGE == GS_AFTER) st.median_after = (double)med;
    else if (STAGE == GS_PROC) st.centre = (double)med;
    else {
        const T c = (T)st.centre, d = med;
        const T ds = d * (T)g.sigma;   // preprocessor.py:739-740, arithmetic in the data's precision
        const T hi = c + ds, lo = c - ds;
        st.mad = (double)d; st.thr_hi = (double)hi; st.thr_lo = (double)lo;
    }
}

// ------------------------------------------------------------------------------------------
// counting passes: +-inf / finite after the stretch; flagged samples
template <int DT>
__global__ void __launch_bounds__(kGT)
ginf_count_kernel(GGeom g, const void* __restrict__ data, rfi_tile_stat_t* __restrict__ stats,
                  GSel* __restrict__ sel) {
    using T = typename In<DT>::T;
    const long long grp = active_group(g, blockIdx.x / g.chunks);
    const int chunk = blockIdx.x % g.chunks;
    const GCtx<T> ctx = load_ctx<T>(g, stats[grp], sel[grp]);
    const GWindow win = group_window(g, grp);
    const long long n = (long long)g.Pr * g.Pc;
    unsigned int ninf = 0, nfin = 0;
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;  // < 2^31 (make_geom)
        if (e >= n) break;
        const uint32_t yy = e / (uint32_t)g.Pc, xx = e - yy * (uint32_t)g.Pc;
        T a, ph;
        load_src<DT, false>(data, g, win.w, win.r0 + yy, win.c0 + xx, a, ph);
        if (ctx.divide) a = a / ctx.m;
        a = apply_stretch<T>(a, g.stretch);
        const bool inf = is_inf(a);
        ninf += inf ? 1u : 0u;
        nfin += (!inf && !is_nan(a)) ? 1u : 0u;
    }
    ninf = warp_sum_u32(ninf);
    nfin = warp_sum_u32(nfin);
    if ((threadIdx.x & 31) == 0) {
        if (ninf) atomicAdd(reinterpret_cast<unsigned int*>(&stats[grp].n_inf), ninf);
        if (nfin) atomicAdd(&sel[grp].n_fin, nfin);
    }
}

template <int DT>
__global__ void __launch_bounds__(kGT)
gflag_count_kernel(GGeom g, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                   rfi_tile_stat_t* __restrict__ stats, const GSel* __restrict__ sel) {
    using T = typename In<DT>::T;
    const long long grp = active_group(g, blockIdx.x / g.chunks);
    const int chunk = blockIdx.x % g.chunks;
    const GCtx<T> ctx = load_ctx<T>(g, stats[grp], sel[grp]);
    const GWindow win = group_window(g, grp);
    const long long n = (long long)g.Pr * g.Pc;
    unsigned int nf = 0;
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;
        if (e >= n) break;
        const uint32_t yy = e / (uint32_t)g.Pc;
        const long long row = win.r0 + yy, col = win.c0 + (e - yy * (uint32_t)g.Pc);
        if (g.flag_mode == RFI_FLAGS_MAD) {
            T a, ph;
            load_src<DT, false>(data, g, win.w, row, col, a, ph);
            const T x = processed<T>(a, g, ctx);
            nf += ((x > ctx.thr_hi) || (x < ctx.thr_lo)) ? 1u : 0u;
        } else if (row >= 0 && row < g.C && col >= 0 && col < g.T) {
            nf += __ldg(flags + (size_t)((win.w * g.C + row) * g.T + col)) != 0 ? 1u : 0u;
        }
    }
    nf = warp_sum_u32(nf);
    if ((threadIdx.x & 31) == 0 && nf) atomicAdd(reinterpret_cast<unsigned int*>(&stats[grp].n_flagged), nf);
}

__global__ void gstats_init_kernel(GGeom g, rfi_tile_stat_t* __restrict__ stats, GSel* __restrict__ sel) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= g.n_active) return;
    const long long grp = active_group(g, gi);
    rfi_tile_stat_t st;
    st.median_before = st.inf_fill = st.median_after = 0.0;
    st.centre = st.mad = st.thr_lo = st.thr_hi = 0.0;
    st.n_valid = g.Pr * g.Pc; st.n_inf = 0; st.n_flagged = 0; st.raw_lo = st.raw_hi = 0.0;
    st.route = g.list ? RFI_TILE_GENERAL : 0;
    stats[grp] = st;
    GSel s;
    s.prefix = 0; s.nxt = ~0ull; s.cle = s.n = s.k1 = s.k2 = 0; s.shift = 0; s.pad = 0;
    s.cf = 0.0; s.n_fin = 0; s.pad2 = 0;
    for (int t = 0; t < 16; ++t) s.cnt[t] = 0;
    sel[grp] = s;
}

// ------------------------------------------------------------------------------------------
// phase 2
template <int DT, bool kPhase>
RFI_DEVINL void eval_pixel(const void* data, const GGeom& g, const GPatch& p, const GCtx<typename In<DT>::T>& ctx,
                           int y, int x, typename In<DT>::T& xproc, typename In<DT>::T& L,
                           typename In<DT>::T& ph) {
    using T = typename In<DT>::T;
    const long long row = p.r0 + (long long)y * p.ry + (long long)x * p.rx;
    const long long col = p.c0 + (long long)y * p.cy + (long long)x * p.cx;
    T a;
    load_src<DT, kPhase>(data, g, p.w, row, col, a, ph);
    xproc = processed<T>(a, g, ctx);
    L = log10_img(fabs_(xproc) + T(1e-10));
}

template <int DT>
RFI_DEVINL typename In<DT>::T eval_L(const void* data, const GGeom& g, const GPatch& p,
                                     const GCtx<typename In<DT>::T>& ctx, int y, int x) {
    using T = typename In<DT>::T;
    T xp, L, ph;
    eval_pixel<DT, false>(data, g, p, ctx, y, x, xp, L, ph);
    return L;
}

__global__ void grange_init_kernel(long long n, GRange* __restrict__ rng) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rng[i].lmin = rng[i].gmin = ~0ull;
    rng[i].lmax = rng[i].gmax = 0ull;  // key 0 is never produced by a non-NaN value
}

template <typename K>
RFI_DEVINL void block_key_range(K lo, K hi, unsigned long long* gmin, unsigned long long* gmax) {
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != ~K(0)) atomicMin(gmin, (unsigned long long)lo);
        if (hi != 0) atomicMax(gmax, (unsigned long long)hi);
    }
}

// blockIdx.x = kept-patch ordinal * chunks + chunk; `kept` lists the canonical indices to write
template <int DT, bool kComplexBranch>
__global__ void __launch_bounds__(kGT)
grange_kernel(GGeom g, const void* __restrict__ data, const rfi_tile_stat_t* __restrict__ stats,
              const GSel* __restrict__ sel, const long long* __restrict__ dest_slot,
              GRange* __restrict__ rng) {
    using T = typename In<DT>::T;
    using K = typename Scalar<T>::key_t;
    const long long q = blockIdx.x / g.chunks;
    const int chunk = blockIdx.x % g.chunks;
    if (dest_slot[q] < 0) return;
    const GPatch p = patch_map(g, q);
    const GCtx<T> ctx = load_ctx<T>(g, stats[p.grp], sel[p.grp]);
    const long long n = (long long)p.rows * p.cols;
    K llo = ~K(0), lhi = 0, glo = ~K(0), ghi = 0;
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;
        if (e >= n) break;
        const int y = (int)(e / (uint32_t)p.cols), x = (int)(e - (uint32_t)y * (uint32_t)p.cols);
        const T c = eval_L<DT>(data, g, p, ctx, y, x);
        const T td = (y > 0) ? c - eval_L<DT>(data, g, p, ctx, y - 1, x) : T(0);
        const T fd = (x > 0) ? c - eval_L<DT>(data, g, p, ctx, y, x - 1) : T(0);
        const T gr = sqrt_fast(td * td + fd * fd);
        const K kl = to_key<T>(c), kg = to_key<T>(gr);
        if (kl != ~K(0)) { llo = kl < llo ? kl : llo; lhi = kl > lhi ? kl : lhi; }
        if (kg != ~K(0)) { glo = kg < glo ? kg : glo; ghi = kg > ghi ? kg : ghi; }
    }
    if constexpr (!kComplexBranch) block_key_range<K>(llo, lhi, &rng[q].lmin, &rng[q].lmax);
    block_key_range<K>(glo, ghi, &rng[q].gmin, &rng[q].gmax);
}

template <typename T>
RFI_DEVINL ChanScale<T> scale_from_keys(unsigned long long kmin, unsigned long long kmax) {
    using K = typename Scalar<T>::key_t;
    const T lo = (kmin == ~0ull) ? Scalar<T>::nan() : from_key<T>((K)kmin);
    const T hi = (kmax == 0ull) ? Scalar<T>::nan() : from_key<T>((K)kmax);
    return make_scale<T>(lo, hi);
}

template <int DT, bool kComplexBranch>
__global__ void __launch_bounds__(kGT)
gwrite_kernel(GGeom g, const void* __restrict__ data, const uint8_t* __restrict__ flags,
              const rfi_tile_stat_t* __restrict__ stats, const GSel* __restrict__ sel,
              const long long* __restrict__ dest_slot, const GRange* __restrict__ rng,
              float* __restrict__ images, uint8_t* __restrict__ labels) {
    using T = typename In<DT>::T;
    const long long q = blockIdx.x / g.chunks;
    const int chunk = blockIdx.x % g.chunks;
    const long long slot = dest_slot[q];
    if (slot < 0) return;
    const GPatch p = patch_map(g, q);
    const GCtx<T> ctx = load_ctx<T>(g, stats[p.grp], sel[p.grp]);
    const long long n = (long long)p.rows * p.cols;
    const GRange r = rng[q];
    const ChanScale<T> gs = scale_from_keys<T>(r.gmin, r.gmax);
    const ChanScale<T> ls = scale_from_keys<T>(r.lmin, r.lmax);
    const float mean2 = 0.406f;
    const float std0 = 0.229f, std1 = 0.224f, std2 = 0.225f;
    const float is0 = 1.0f / std0, is1 = 1.0f / std1;
    const float nb0 = (0.0f - 0.485f) / std0, nb1 = (0.0f - 0.456f) / std1, nb2 = (0.0f - mean2) / std2;
    float* out_img = images + (size_t)slot * n * 3;
    uint8_t* out_lab = labels + (size_t)slot * n;
    for (int k = 0; k < kGE; ++k) {
        const uint32_t e = (uint32_t)chunk * kGChunk + (uint32_t)k * kGT + threadIdx.x;
        if (e >= n) break;
        const int y = (int)(e / (uint32_t)p.cols), x = (int)(e - (uint32_t)y * (uint32_t)p.cols);
        T xp, c, ph;
        eval_pixel<DT, kComplexBranch>(data, g, p, ctx, y, x, xp, c, ph);
        const T td = (y > 0) ? c - eval_L<DT>(data, g, p, ctx, y - 1, x) : T(0);
        const T fd = (x > 0) ? c - eval_L<DT>(data, g, p, ctx, y, x - 1) : T(0);
        const T gr = sqrt_fast(td * td + fd * fd);
        const float u0 = gs.ok ? (float)((gr - gs.lo) * gs.inv) : 0.f;   // flat / all-NaN channel: zeros
        float o1, o2;
        if constexpr (kComplexBranch) {
            T u = (c - T(-3.0)) * T(1.0 / 7.0);
            u = u < T(0) ? T(0) : (u > T(1) ? T(1) : u);
            o1 = __fmaf_rn((float)u, is1, nb1);
            const T c2 = (ph + T(3.141592653589793)) / T(6.283185307179586);
            o2 = ((float)c2 - mean2) / std2;
        } else {
            const float u1 = ls.ok ? (float)((c - ls.lo) * ls.inv) : 0.f;
            o1 = __fmaf_rn(u1, is1, nb1);
            o2 = nb2;
        }
        out_img[e * 3 + 0] = __fmaf_rn(u0, is0, nb0);
        out_img[e * 3 + 1] = o1;
        out_img[e * 3 + 2] = o2;
        unsigned char f = 0;
        if (g.flag_mode == RFI_FLAGS_MAD) {
            f = ((xp > ctx.thr_hi) || (xp < ctx.thr_lo)) ? 1 : 0;
        } else if (g.flag_mode == RFI_FLAGS_CUSTOM) {
            const long long row = p.r0 + (long long)y * p.ry + (long long)x * p.rx;
            const long long col = p.c0 + (long long)y * p.cy + (long long)x * p.cx;
            if (row >= 0 && row < g.C && col >= 0 && col < g.T)
                f = __ldg(flags + (size_t)((p.w * g.C + row) * g.T + col));
        }
        out_lab[e] = f;
    }
}

// ------------------------------------------------------------------------------------------
// host side
static int make_geom(const rfi_plan_t* plan, GGeom& g) {
    if (!plan) { set_error("plan is NULL"); return RFI_E_INVALID; }
    if (plan->dtype < RFI_F32 || plan->dtype > RFI_C128) { set_error("bad dtype %d", plan->dtype); return RFI_E_INVALID; }
    if (plan->rotations != 1 && plan->rotations != 2 && plan->rotations != 4) {
        set_error("rotations must be 1, 2 or 4 (got %d)", plan->rotations); return RFI_E_INVALID; }
    if (plan->stretch < 0 || plan->stretch > 2) { set_error("bad stretch %d", plan->stretch); return RFI_E_INVALID; }
    if (plan->flag_mode < 0 || plan->flag_mode > 2) { set_error("bad flag_mode %d", plan->flag_mode); return RFI_E_INVALID; }
    if (plan->n_waterfalls < 0 || plan->channels <= 0 || plan->times <= 0) { set_error("bad cube shape"); return RFI_E_INVALID; }
    if (plan->patch <= 0) { set_error("bad patch size %d", plan->patch); return RFI_E_INVALID; }
    const long long C = plan->channels, T = plan->times;
    const int P = plan->patch;
    g.C = C; g.T = T; g.n_waterfalls = plan->n_waterfalls; g.P = P; g.R = plan->rotations;
    g.skip = (C <= P && T <= P) ? 1 : 0;  // preprocessor.py:261
    if (g.skip) {
        if (g.R == 4 && C != T) {
            set_error("rotated views of a non-square %lld x %lld waterfall cannot be stacked when patchify is skipped "
                      "(the reference raises here too)", C, T);
            return RFI_E_INVALID;
        }
        g.Pr = (int)C; g.Pc = (int)T; g.nhc = g.nwc = 1; g.padded = 0;
    } else {
        g.Pr = g.Pc = P;
        g.nhc = (int)((C + P - 1) / P); g.nwc = (int)((T + P - 1) / P);
        g.padded = (C % P || T % P) ? 1 : 0;
    }
    g.per = g.nhc * g.nwc;
    g.n_patches = g.n_waterfalls * g.R * g.per;
    g.n_groups = g.padded ? g.n_patches : g.n_waterfalls * g.per;
    const long long n = (long long)g.Pr * g.Pc;
    if (n > 0x40000000LL) { set_error("patch of %lld samples is too large", n); return RFI_E_UNSUPPORTED; }
    g.chunks = (int)((n + kGChunk - 1) / kGChunk);
    g.stretch = plan->stretch; g.norm_before = plan->norm_before; g.norm_after = plan->norm_after;
    g.flag_mode = plan->flag_mode; g.sigma = plan->sigma;
    g.real_branch = (plan->dtype < RFI_C64 || plan->magnitude) ? 1 : 0;
    g.list = nullptr; g.n_active = g.n_groups;
    if (!g.real_branch) { g.stretch = RFI_STRETCH_NONE; g.norm_before = g.norm_after = 0; }
    if ((long long)g.chunks * (g.n_patches > g.n_groups ? g.n_patches : g.n_groups) > 0x7fffffffLL) {
        set_error("cube too large for one launch of the generic path"); return RFI_E_UNSUPPORTED; }
    return RFI_OK;
}

bool plan_is_fast(const rfi_plan_t* plan) {
    return plan && plan->patch == kP && plan->channels >= kP && plan->times >= kP &&
           plan->channels % kP == 0 && plan->times % kP == 0;
}

long long generic_num_groups(const rfi_plan_t* plan) {
    GGeom g;
    return make_geom(plan, g) ? -1 : g.n_groups;
}
long long generic_num_patches(const rfi_plan_t* plan) {
    GGeom g;
    return make_geom(plan, g) ? -1 : g.n_patches;
}
size_t generic_workspace_bytes(const rfi_plan_t* plan) {
    GGeom g;
    if (make_geom(plan, g)) return 0;
    return (size_t)g.n_groups * sizeof(GSel) + (size_t)g.n_patches * sizeof(GRange) + 256;
}

static GSel* ws_sel(void* ws) { return static_cast<GSel*>(ws); }
static GRange* ws_range(void* ws, const GGeom& g) {
    size_t off = ((size_t)g.n_groups * sizeof(GSel) + 255) & ~(size_t)255;
    return reinterpret_cast<GRange*>(static_cast<char*>(ws) + off);
}

template <int DT, int STAGE>
static void run_select(const GGeom& g, const void* data, rfi_tile_stat_t* stats, GSel* sel, cudaStream_t st) {
    using T = typename In<DT>::T;
    constexpr int kBits = sizeof(T) * 8;
    const unsigned gthreads = 128, ggrid = (unsigned)((g.n_active + gthreads - 1) / gthreads);
    const unsigned grid = (unsigned)(g.n_active * g.chunks);
    gsel_begin_kernel<<<ggrid, gthreads, 0, st>>>(g, sel, kBits);
    for (int p = 0; p < kBits / 4; ++p) {
        gsel_count_kernel<DT, STAGE><<<grid, kGT, 0, st>>>(g, data, stats, sel);
        gsel_decide_kernel<<<ggrid, gthreads, 0, st>>>(g, sel, p == 0 ? 1 : 0);
    }
    gsel_next_kernel<DT, STAGE><<<grid, kGT, 0, st>>>(g, data, stats, sel);
    gsel_finish_kernel<T, STAGE><<<ggrid, gthreads, 0, st>>>(g, stats, sel);
}

template <int DT>
static int run_generic_stats(const GGeom& g, const void* data, const uint8_t* flags,
                             rfi_tile_stat_t* stats, GSel* sel, cudaStream_t st) {
    const unsigned gthreads = 128, ggrid = (unsigned)((g.n_active + gthreads - 1) / gthreads);
    const unsigned grid = (unsigned)(g.n_active * g.chunks);
    gstats_init_kernel<<<ggrid, gthreads, 0, st>>>(g, stats, sel);
    if (g.real_branch) {
        if (g.norm_before) run_select<DT, GS_RAW>(g, data, stats, sel, st);
        if (g.stretch != RFI_STRETCH_NONE) {
            ginf_count_kernel<DT><<<grid, kGT, 0, st>>>(g, data, stats, sel);
            run_select<DT, GS_FIN>(g, data, stats, sel, st);
            run_select<DT, GS_FIN_DEV>(g, data, stats, sel, st);
        }
        if (g.norm_after) run_select<DT, GS_AFTER>(g, data, stats, sel, st);
    }
    if (g.flag_mode == RFI_FLAGS_MAD) {
        run_select<DT, GS_PROC>(g, data, stats, sel, st);
        run_select<DT, GS_PROC_DEV>(g, data, stats, sel, st);
        gflag_count_kernel<DT><<<grid, kGT, 0, st>>>(g, data, flags, stats, sel);
    } else if (g.flag_mode == RFI_FLAGS_CUSTOM) {
        gflag_count_kernel<DT><<<grid, kGT, 0, st>>>(g, data, flags, stats, sel);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

int generic_tile_stats(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                       rfi_tile_stat_t* stats, void* workspace, cudaStream_t st) {
    GGeom g;
    int rc = make_geom(plan, g);
    if (rc) return rc;
    if (g.n_groups == 0) return RFI_OK;
    if (!data || !stats) { set_error("data / stats is NULL"); return RFI_E_INVALID; }
    if (!workspace) { set_error("this geometry runs on the generic path and needs a workspace of "
                                "rfi_plan_workspace_bytes() bytes"); return RFI_E_INVALID; }
    if (g.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    GSel* sel = ws_sel(workspace);
    switch (plan->dtype) {
        case RFI_F32: return run_generic_stats<RFI_F32>(g, data, flags, stats, sel, st);
        case RFI_F64: return run_generic_stats<RFI_F64>(g, data, flags, stats, sel, st);
        case RFI_C64: return run_generic_stats<RFI_C64>(g, data, flags, stats, sel, st);
        default:      return run_generic_stats<RFI_C128>(g, data, flags, stats, sel, st);
    }
}

int generic_tile_stats_subset(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                              rfi_tile_stat_t* stats, void* workspace, const int* list, int n_list,
                              cudaStream_t st) {
    GGeom g;
    int rc = make_geom(plan, g);
    if (rc) return rc;
    if (n_list <= 0) return RFI_OK;
    if (g.padded || g.skip) { set_error("internal: group subsets need an unpadded geometry"); return RFI_E_INVALID; }
    g.list = list; g.n_active = n_list;
    GSel* sel = ws_sel(workspace);
    switch (plan->dtype) {
        case RFI_F32: return run_generic_stats<RFI_F32>(g, data, flags, stats, sel, st);
        case RFI_F64: return run_generic_stats<RFI_F64>(g, data, flags, stats, sel, st);
        case RFI_C64: return run_generic_stats<RFI_C64>(g, data, flags, stats, sel, st);
        default:      return run_generic_stats<RFI_C128>(g, data, flags, stats, sel, st);
    }
}

template <int DT, bool CB>
static int run_generic_write(const GGeom& g, const void* data, const uint8_t* flags,
                             const rfi_tile_stat_t* stats, const GSel* sel, const long long* dest,
                             GRange* rng, float* images, uint8_t* labels, cudaStream_t st) {
    const unsigned grid = (unsigned)(g.n_patches * g.chunks);
    grange_init_kernel<<<(unsigned)((g.n_patches + 255) / 256), 256, 0, st>>>(g.n_patches, rng);
    grange_kernel<DT, CB><<<grid, kGT, 0, st>>>(g, data, stats, sel, dest, rng);
    gwrite_kernel<DT, CB><<<grid, kGT, 0, st>>>(g, data, flags, stats, sel, dest, rng, images, labels);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

int generic_write_patches(const rfi_plan_t* plan, const void* data, const uint8_t* flags,
                          const rfi_tile_stat_t* stats, const long long* dest_slot, float* images,
                          uint8_t* labels, void* workspace, cudaStream_t st) {
    GGeom g;
    int rc = make_geom(plan, g);
    if (rc) return rc;
    if (g.n_patches == 0) return RFI_OK;
    if (!data || !stats || !dest_slot) { set_error("data / stats / dest_slot is NULL"); return RFI_E_INVALID; }
    if (!workspace) { set_error("this geometry runs on the generic path and needs the workspace "
                                "rfi_tile_stats was given"); return RFI_E_INVALID; }
    if (g.flag_mode == RFI_FLAGS_CUSTOM && !flags) { set_error("custom flag mode needs flags"); return RFI_E_INVALID; }
    const GSel* sel = ws_sel(workspace);
    GRange* rng = ws_range(workspace, g);
    const bool cb = !g.real_branch;
    switch (plan->dtype) {
        case RFI_F32: return run_generic_write<RFI_F32, false>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st);
        case RFI_F64: return run_generic_write<RFI_F64, false>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st);
        case RFI_C64:
            return cb ? run_generic_write<RFI_C64, true>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st)
                      : run_generic_write<RFI_C64, false>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st);
        default:
            return cb ? run_generic_write<RFI_C128, true>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st)
                      : run_generic_write<RFI_C128, false>(g, data, flags, stats, sel, dest_slot, rng, images, labels, st);
    }
}

}  // namespace rfi
