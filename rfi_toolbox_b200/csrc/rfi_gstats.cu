// rfi_gstats.cu -- compute_statistics / compute_ffi over a whole cube (float32 arithmetic) in THREE passes
// over the data instead of twenty.
//
// Replaces rfi_toolbox/evaluation/statistics.py:10-97 for float32 / complex64 input: statistics of ALL
// samples and of the UNFLAGGED samples (`data[~flags]`) -- mean, population std, median, MAD, max, counts
// -- from which compute_ffi forms its reductions.  (float64 / complex128 keep the radix passes of
// rfi_stats.cu.)
//
//   pass A   the cube once (8 B / px + 1 B of flags): NumPy-exact |z| -> order-preserving keys left in a
//            4 B / px scratch; counts, NaN counts, maxima and float64 sums for both sets
//   pass B   the scratch (4 + 1 B / px): squared deviations from the two means (np.std's own formula); for
//            each set, how many keys lie below a bracket [lo, hi] that holds the median with overwhelming
//            probability, and the keys inside it (2.5 % of the data) compacted into a list
//   pass C   the scratch again: the same for the keys |x - median| (the MAD)
// A bracket comes from a stratified random sample of 8192 keys (rank +- 4.5 sigma), the exact rank
// inside a list of up to 32768 keys from one CTA (mono_resolve: linear histograms refined on the
// answer's bucket); a longer list goes through one more sample / bracket / compaction level first
// (20 x shorter each time), reading only the list -- the number of levels follows from n on the host.  Everything is stream-ordered with device-resident
// sizes; a bracket that misses its rank (probability ~1e-5 per call) or a list that overflows is reported
// (count = -1) and the caller repeats the call through the 4-bit radix passes of rfi_stats.cu: results are
// exact order statistics either way.
#include "rfi_stats_mono.cuh"

namespace rfi {

constexpr int kGsSample = 8192;       // keys per sample (a bracket then keeps ~5 % of its source)
constexpr int kGsFinal = 32768;       // a list up to this long is resolved by one CTA
constexpr int kGsMaxLevels = 5;       // compaction levels (20 x shorter each): cubes up to ~1e11 samples
constexpr float kGsSigma = 4.5f;      // half width of a bracket in sigma of a sample rank
constexpr int kGsNT = 256;            // threads of the streaming kernels
constexpr int kGsChunk = 2048;        // keys per CTA iteration of the compaction
constexpr int kGsStage = 2 * kGsChunk;  // staged candidates per set: flushed (ONE global atomic) when a further
                                        // chunk might not fit -- every ~20 chunks at 5 %, every chunk at 100 %
using GK = uint32_t;
constexpr GK kGExcl = ~GK(0);

struct GsSel {                 // selection state of one set
    unsigned long long n_valid;    // valid (non-NaN) keys of the set
    unsigned long long k1, k2;     // target ranks (0-based) among them
    unsigned long long below;      // keys proven below the current list
    unsigned long long m;          // keys in the current source (level 0: n_valid; then the list sizes)
    unsigned long long m_next;     // cursor of the list being built
    GK lo, hi;                     // bracket of the compaction in flight
    GK key1, key2;                 // the two order statistics
    int skip, pad;                 // nothing to select (no valid key; MAD of a non-finite median)
};
struct GsWork {
    unsigned long long n[2], n_nan[2], n_flagged;
    double sum[2], sumsq[2];       // sum of x; sum of (x - mean)^2
    GK maxkey[2];
    float mean[2], centre[2], median[2], mad[2];
    int fail;                      // a bracket missed / a list overflowed: the caller falls back to the radix passes
    int dev_mode;                  // 0: keys are magnitudes; 1: keys are |x - centre|
    GsSel sel[2];
};

// scratch key of an element for set s at level 0 (flagged -> excluded for set 1; deviation in dev mode)
RFI_DEVINL GK gs_key0(GK k, bool flagged, int set, int dev, float centre) {
    if (k == kGExcl || (set == 1 && flagged)) return kGExcl;
    if (dev) return to_key<float>(fabsf(from_key<float>(k) - centre));
    return k;
}

RFI_DEVINL unsigned long long gs_hash(unsigned long long x) {   // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

// block totals of a few doubles / 64-bit counts -> thread 0 (kGsNT threads)
template <int ND, int NU>
RFI_DEVINL void gs_block_totals(double (&d)[ND], unsigned long long (&u)[NU]) {
    __shared__ double sd[ND > 0 ? ND : 1][kGsNT / 32];
    __shared__ unsigned long long su[NU > 0 ? NU : 1][kGsNT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int o = 16; o; o >>= 1) d[i] += __shfl_xor_sync(0xffffffffu, d[i], o);
#pragma unroll
    for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int o = 16; o; o >>= 1) u[i] += __shfl_xor_sync(0xffffffffu, u[i], o);
    __syncthreads();
    if (lane == 0) {
        for (int i = 0; i < ND; ++i) sd[i][warp] = d[i];
        for (int i = 0; i < NU; ++i) su[i][warp] = u[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < ND; ++i) { double a = 0.0; for (int v = 0; v < kGsNT / 32; ++v) a += sd[i][v]; d[i] = a; }
        for (int i = 0; i < NU; ++i) { unsigned long long a = 0; for (int v = 0; v < kGsNT / 32; ++v) a += su[i][v]; u[i] = a; }
    }
}

// ---- pass A: cube -> keys in the scratch, sums, counts, maxima -------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kGsNT)
gs_pass_a_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long n,
                 GK* __restrict__ keys, GsWork* __restrict__ w) {
    double s[2] = {0.0, 0.0};
    unsigned long long cnt[3] = {0, 0, 0};   // NaNs of the two sets, flagged samples
    GK mx[2] = {0, 0};
    const long long groups = n >> 2;
    for (long long g = (long long)blockIdx.x * kGsNT + threadIdx.x; g <= groups; g += (long long)gridDim.x * kGsNT) {
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t f4 = 0;
        int nin = 0;
        const long long i0 = g << 2;
        if (g < groups) {
            load4_mag_fast<DT>(data, (size_t)i0, x);
            if (flags) f4 = __ldg(reinterpret_cast<const uint32_t*>(flags + i0));
            nin = 4;
        } else {  // tail of fewer than four samples
            for (int j = 0; i0 + j < n; ++j) {
                if constexpr (DT == RFI_C64) { const float2 z = __ldg(static_cast<const float2*>(data) + i0 + j); x[j] = cabs_np<float>(z.x, z.y); }
                else x[j] = __ldg(static_cast<const float*>(data) + i0 + j);
                if (flags) f4 |= (uint32_t)__ldg(flags + i0 + j) << (8 * j);
                nin = j + 1;
            }
            if (nin == 0) continue;
        }
        GK k4[4];
        float sa = 0.f, sc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool in = j < nin, fl = ((f4 >> (8 * j)) & 0xffu) != 0;
            const GK k = to_key<float>(x[j]);
            k4[j] = k;
            const bool nan = in && k == kGExcl;
            cnt[0] += nan ? 1u : 0u;
            cnt[1] += (nan && !fl) ? 1u : 0u;
            cnt[2] += (in && fl) ? 1u : 0u;
            const GK km = in ? k : GK(0);
            mx[0] = km > mx[0] ? km : mx[0];
            const GK kc = fl ? GK(0) : km;
            mx[1] = kc > mx[1] ? kc : mx[1];
            sa += in ? x[j] : 0.f;
            sc += (in && !fl) ? x[j] : 0.f;
        }
        s[0] += (double)sa; s[1] += (double)sc;
        if (nin == 4) *reinterpret_cast<uint4*>(keys + i0) = make_uint4(k4[0], k4[1], k4[2], k4[3]);
        else for (int j = 0; j < nin; ++j) keys[i0 + j] = k4[j];
    }
    __shared__ GK sm[2][kGsNT / 32];
    mx[0] = warp_max(mx[0]); mx[1] = warp_max(mx[1]);
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = mx[0]; sm[1][threadIdx.x >> 5] = mx[1]; }
    gs_block_totals<2, 3>(s, cnt);   // (its barriers publish sm)
    if (threadIdx.x == 0) {
        GK c[2] = {0, 0};
        for (int v = 0; v < kGsNT / 32; ++v) { c[0] = sm[0][v] > c[0] ? sm[0][v] : c[0]; c[1] = sm[1][v] > c[1] ? sm[1][v] : c[1]; }
        atomicAdd(&w->sum[0], s[0]); atomicAdd(&w->sum[1], s[1]);
        if (cnt[0]) atomicAdd(&w->n_nan[0], cnt[0]);
        if (cnt[1]) atomicAdd(&w->n_nan[1], cnt[1]);
        if (cnt[2]) atomicAdd(&w->n_flagged, cnt[2]);
        if (c[0]) atomicMax(&w->maxkey[0], c[0]);
        if (c[1]) atomicMax(&w->maxkey[1], c[1]);
    }
}

// counts, means, selection targets of the median phase
__global__ void gs_after_a_kernel(long long n, GsWork* w) {
    w->n[0] = (unsigned long long)n;
    w->n[1] = (unsigned long long)n - w->n_flagged;
    for (int s = 0; s < 2; ++s) {
        w->mean[s] = w->n[s] ? (float)(w->sum[s] / (double)w->n[s]) : 0.f;
        GsSel& e = w->sel[s];
        e.n_valid = w->n[s] - w->n_nan[s];
        e.k1 = e.n_valid ? (e.n_valid - 1) >> 1 : 0;
        e.k2 = e.n_valid >> 1;
        e.below = 0; e.m = e.n_valid; e.m_next = 0;
        e.skip = e.n_valid == 0;
        w->centre[s] = 0.f;
    }
    w->dev_mode = 0;
}

// ---- sample of a source: level 0 = the scratch (with flags / deviation), level >= 1 = a list ------------------
__global__ void __launch_bounds__(256)
gs_sample_kernel(const GK* __restrict__ src0, const uint8_t* __restrict__ flags, long long n0,
                 const GK* __restrict__ list, unsigned long long list_stride, int level,
                 const GsWork* __restrict__ w, GK* __restrict__ sample) {
    const int set = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;    // sample slot
    const unsigned long long total = level == 0 ? (unsigned long long)n0 : w->sel[set].m;
    GK out = kGExcl;
    if (total > 0 && !w->sel[set].skip) {
        const unsigned long long stride = total / kGsSample > 0 ? total / kGsSample : 1;
        const unsigned long long pos = (unsigned long long)j * stride +
            (stride > 1 ? gs_hash(((unsigned long long)(level * 2 + set) << 40) ^ (unsigned long long)j) % stride : 0);
        if (pos < total) {
            if (level == 0) out = gs_key0(src0[pos], flags && flags[pos] != 0, set, w->dev_mode, w->centre[set]);
            else out = list[(size_t)set * list_stride + pos];
        }
    }
    sample[(size_t)set * kGsSample + j] = out;
}

// ---- bracket of the next compaction from the sample: CTA 2 * set + which finds the lower / upper end ----------
__global__ void __launch_bounds__(512)
gs_bracket_kernel(const GK* __restrict__ sample, GsWork* __restrict__ w) {
    __shared__ MonoShared<GK> sh;
    __shared__ uint32_t cnt;
    const int set = blockIdx.x >> 1, upper = blockIdx.x & 1;
    const GK* smp = sample + (size_t)set * kGsSample;
    GsSel& e = w->sel[set];
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    uint32_t c = 0;
    for (int i = threadIdx.x; i < kGsSample; i += 512) c += smp[i] != kGExcl ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&cnt, c);
    __syncthreads();
    const uint32_t sv = cnt;
    const unsigned long long N = e.m;
    GK end = upper ? kGExcl - 1 : GK(0);
    if (!e.skip && N > (unsigned long long)kGsFinal && sv >= 1024 && e.k1 >= e.below) {
        const unsigned long long kk = e.k1 - e.below;          // target rank inside the source
        const double rho = ((double)(2 * kk + 1) * (double)sv) / (double)(2 * N);
        const int delta = (int)(kGsSigma * 0.5f * sqrtf((float)sv)) + 2;
        const long long at = upper ? (long long)rho + delta + 1 : (long long)rho - delta;
        GK a, b;
        if (at >= 0 && at < (long long)sv) { mono_resolve<GK, 512>(smp, kGsSample, (uint32_t)at, (uint32_t)at, a, b, sh); end = a; }
    }
    if (threadIdx.x == 0) {
        if (upper) e.hi = end; else { e.lo = end; e.m_next = 0; }
    }
}

// ---- compaction: keys below the bracket counted, keys inside it appended to the next list --------------------
// LEVEL0: one pass over the scratch serves both sets (and, MOMENTS, their squared deviations from the
// means); else blockIdx.y = set and the source is a list
template <bool LEVEL0, bool MOMENTS>
__global__ void __launch_bounds__(kGsNT)
gs_compact_kernel(const GK* __restrict__ src0, const uint8_t* __restrict__ flags, long long n0,
                  const GK* __restrict__ in_list, unsigned long long in_stride,
                  GK* __restrict__ out_list, unsigned long long out_stride, unsigned long long cap,
                  GsWork* __restrict__ w) {
    __shared__ GK stage[2][kGsStage];
    __shared__ uint32_t cur[2];
    __shared__ unsigned long long gbase[2];
    const int tid = threadIdx.x;
    const int set0 = LEVEL0 ? 0 : blockIdx.y, nset = LEVEL0 ? 2 : 1;
    const unsigned long long total = LEVEL0 ? (unsigned long long)n0 : w->sel[set0].m;
    const int dev = w->dev_mode;
    GK lo[2], span[2];
    float centre[2], mean[2];
    bool skip[2];
    for (int s = 0; s < 2; ++s) {
        lo[s] = w->sel[s].lo; span[s] = w->sel[s].hi - w->sel[s].lo; centre[s] = w->centre[s]; mean[s] = w->mean[s];
        skip[s] = w->sel[s].skip != 0;
    }
    unsigned long long below[2] = {0, 0};
    double q[2] = {0.0, 0.0};
    const unsigned long long nchunks = (total + kGsChunk - 1) / kGsChunk;
    // staged candidates -> the list: one global atomic per set reserves the range (uniform control flow)
    auto flush = [&](bool want0, bool want1) {
        if (tid < nset && (tid == 0 ? want0 : want1)) {
            const int s = LEVEL0 ? tid : set0;
            const uint32_t c = cur[tid];
            unsigned long long b = ~0ull;
            if (c) {
                b = atomicAdd(&w->sel[s].m_next, (unsigned long long)c);
                if (b + c > cap) { w->fail = 1; b = ~0ull; }
            }
            gbase[tid] = b;
        }
        __syncthreads();
        for (int t = 0; t < nset; ++t) {
            if (!(t == 0 ? want0 : want1)) continue;
            const int s = LEVEL0 ? t : set0;
            const unsigned long long b = gbase[t];
            if (b != ~0ull) {
                const uint32_t c = cur[t];
                for (uint32_t i = tid; i < c; i += kGsNT) out_list[(size_t)s * out_stride + b + i] = stage[t][i];
            }
        }
        __syncthreads();
        if (tid < nset && (tid == 0 ? want0 : want1)) cur[tid] = 0;
        __syncthreads();
    };
    if (tid < 2) cur[tid] = 0;
    for (unsigned long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        __syncthreads();   // the previous chunk is staged
        {
            const bool f0 = cur[0] > (uint32_t)(kGsStage - kGsChunk), f1 = nset > 1 && cur[1] > (uint32_t)(kGsStage - kGsChunk);
            if (f0 || f1) flush(f0, f1);
        }
        const unsigned long long c0 = ch * kGsChunk;
        float qa = 0.f, qc = 0.f;
        if (LEVEL0) {
            // the thread's keys of the chunk stay in registers: one sweep counts (per set: keys below the
            // bracket, a bit mask of the keys inside it), one warp scan + ONE shared atomic per warp and set
            // reserves the staging slots, the masks are then walked -- no per-candidate atomics
            constexpr int E = kGsChunk / kGsNT;   // 8 keys per thread: two groups of four consecutive keys
            static_assert(E == 8, "two 128-bit key loads per thread and chunk");
            GK k[E];
            uint32_t f4[2] = {0u, 0u};
            // both groups are loaded before anything is used (128-bit keys + 32-bit flags each)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned long long i0 = c0 + ((unsigned long long)h * kGsNT + tid) * 4;
                if (i0 + 4 <= total) {
                    const uint4 kq = *reinterpret_cast<const uint4*>(src0 + i0);
                    k[h * 4 + 0] = kq.x; k[h * 4 + 1] = kq.y; k[h * 4 + 2] = kq.z; k[h * 4 + 3] = kq.w;
                    if (flags) f4[h] = *reinterpret_cast<const uint32_t*>(flags + i0);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool in = i0 + j < total;
                        k[h * 4 + j] = in ? src0[i0 + j] : kGExcl;
                        if (in && flags) f4[h] |= (uint32_t)flags[i0 + j] << (8 * j);
                    }
                }
            }
            uint32_t flbits = 0;
#pragma unroll
            for (int it = 0; it < E; ++it) {
                const unsigned long long i = c0 + ((unsigned long long)(it >> 2) * kGsNT + tid) * 4 + (it & 3);
                const bool in = i < total;
                const bool fl = ((f4[it >> 2] >> (8 * (it & 3))) & 0xffu) != 0;
                flbits |= (fl ? 1u : 0u) << it;
                if (MOMENTS && in) {   // np.std: abs(x - mean) ** 2 in T, then summed (NaN keys decode to NaN)
                    const float x = from_key<float>(k[it]);
                    const float da = x - mean[0], dc = x - mean[1];
                    qa += da * da;
                    qc += fl ? 0.f : dc * dc;
                }
            }
            const int lane = tid & 31;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t mask = 0, nb = 0;
#pragma unroll
                for (int it = 0; it < E; ++it) {
                    const GK x = skip[s] ? kGExcl : gs_key0(k[it], (flbits >> it) & 1u, s, dev, centre[s]);
                    nb += x < lo[s] ? 1u : 0u;
                    mask |= (((GK)(x - lo[s]) <= span[s] && x != kGExcl) ? 1u : 0u) << it;
                }
                below[s] += nb;
                const uint32_t mine = __popc(mask);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                uint32_t wbase = 0;
                if (lane == 31 && incl) wbase = atomicAdd(&cur[s], incl);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                uint32_t at = wbase + incl - mine;
                while (mask) {
                    const int it = __ffs(mask) - 1;
                    mask &= mask - 1;
                    GK kk = k[0];
#pragma unroll
                    for (int j = 1; j < E; ++j) kk = (j == it) ? k[j] : kk;
                    stage[s][at++] = gs_key0(kk, (flbits >> it) & 1u, s, dev, centre[s]);
                }
            }
        } else {
#pragma unroll 4
            for (int it = 0; it < kGsChunk / kGsNT; ++it) {
                const unsigned long long i = c0 + (unsigned long long)it * kGsNT + tid;
                if (i < total) {
                    const GK x = in_list[(size_t)set0 * in_stride + i];
                    below[0] += x < lo[set0] ? 1u : 0u;
                    if ((GK)(x - lo[set0]) <= span[set0] && x != kGExcl) {
                        const uint32_t at = atomicAdd(&cur[0], 1u);
                        if (at < (uint32_t)kGsStage) stage[0][at] = x;
                    }
                }
            }
        }
        q[0] += (double)qa; q[1] += (double)qc;
    }
    __syncthreads();
    flush(true, true);
    gs_block_totals<2, 2>(q, below);
    if (tid == 0) {
        if (LEVEL0) {
            if (below[0]) atomicAdd(&w->sel[0].below, below[0]);
            if (below[1]) atomicAdd(&w->sel[1].below, below[1]);
            if (MOMENTS) { atomicAdd(&w->sumsq[0], q[0]); atomicAdd(&w->sumsq[1], q[1]); }
        } else if (below[0]) {
            atomicAdd(&w->sel[set0].below, below[0]);
        }
    }
}

// the list just built becomes the source; the target rank must lie inside it
__global__ void gs_advance_kernel(GsWork* w) {
    for (int s = 0; s < 2; ++s) {
        GsSel& e = w->sel[s];
        if (e.skip) continue;
        e.m = e.m_next;
        if (e.below > e.k1 || e.k2 >= e.below + e.m) w->fail = 1;
    }
}

// exact ranks inside the last list (one CTA per set)
__global__ void __launch_bounds__(512)
gs_resolve_kernel(const GK* __restrict__ list, unsigned long long stride, GsWork* __restrict__ w) {
    __shared__ MonoShared<GK> sh;
    const int set = blockIdx.x;
    GsSel& e = w->sel[set];
    if (e.skip || w->fail) return;
    if (e.m == 0 || e.m > (unsigned long long)kGsFinal || e.below > e.k1 || e.k2 >= e.below + e.m) {
        if (threadIdx.x == 0) w->fail = 1;
        return;
    }
    GK a, b;
    mono_resolve<GK, 512>(list + (size_t)set * stride, (uint32_t)e.m, (uint32_t)(e.k1 - e.below), (uint32_t)(e.k2 - e.below), a, b, sh);
    if (threadIdx.x == 0) { e.key1 = a; e.key2 = b; }
}

// median found: it becomes the centre of the deviation keys; the selection restarts for the MAD
__global__ void gs_after_median_kernel(GsWork* w) {
    for (int s = 0; s < 2; ++s) {
        GsSel& e = w->sel[s];
        float med = Scalar<float>::nan();
        if (!e.skip) med = median_of_pair<float>(from_key<float>(e.key1), from_key<float>(e.key2), (uint32_t)(e.n_valid & 1u));
        w->median[s] = med;
        w->centre[s] = med;
        w->mad[s] = Scalar<float>::nan();
        e.skip = e.skip || is_nan(med) || is_inf(med);   // |x - inf| holds a NaN: np.median propagates it
        e.below = 0; e.m = e.n_valid; e.m_next = 0;
    }
    w->dev_mode = 1;
}

__global__ void gs_finish_kernel(GsWork* w, rfi_stats_t* out) {
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    for (int s = 0; s < 2; ++s) {
        GsSel& e = w->sel[s];
        if (!e.skip) w->mad[s] = median_of_pair<float>(from_key<float>(e.key1), from_key<float>(e.key2), (uint32_t)(e.n_valid & 1u));
        rfi_stats_t o;
        o.count = (long long)w->n[s];
        o.n_flagged = s == 1 ? (long long)w->n_flagged : 0;
        o.n_nan = (long long)w->n_nan[s];
        o.mean = o.std = o.median = o.mad = o.max = kNaN;
        if (w->n[s]) {
            o.mean = (double)w->mean[s];
            o.std = (double)__fsqrt_rn((float)(w->sumsq[s] / (double)w->n[s]));
            o.max = (double)from_key<float>(w->maxkey[s]);
            if (w->n_nan[s] == 0) { o.median = (double)w->median[s]; o.mad = (double)w->mad[s]; }
        }
        if (w->fail) o.count = -1;   // a bracket missed: the caller repeats the call through rfi_statistics
        out[s] = o;
    }
}

struct GsLayout {
    size_t work, keys, sample, list1, list2, total;
    unsigned long long cap1, cap2;
};
static GsLayout gs_layout(long long n) {
    GsLayout L;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // a level keeps ~5 % of its source: lists hold 12.5 % (level 0, 2, 4 -> list1; level 1, 3 -> list2)
    L.cap1 = (unsigned long long)(n / 8 > kGsFinal ? n / 8 : kGsFinal);
    L.cap2 = L.cap1 / 8 > (unsigned long long)kGsFinal ? L.cap1 / 8 : (unsigned long long)kGsFinal;
    L.work = 0;
    L.keys = up(sizeof(GsWork));
    L.sample = L.keys + up((size_t)(n > 0 ? n : 1) * 4);
    L.list1 = L.sample + up((size_t)2 * kGsSample * 4);
    L.list2 = L.list1 + up((size_t)2 * L.cap1 * 4);
    L.total = L.list2 + up((size_t)2 * L.cap2 * 4);
    return L;
}

template <int DT>
static int gs_run(const void* data, const uint8_t* flags, long long n, rfi_stats_t* out, void* workspace, cudaStream_t st) {
    const GsLayout L = gs_layout(n);
    char* base = static_cast<char*>(workspace);
    GsWork* w = reinterpret_cast<GsWork*>(base + L.work);
    GK* keys = reinterpret_cast<GK*>(base + L.keys);
    GK* sample = reinterpret_cast<GK*>(base + L.sample);
    GK* list1 = reinterpret_cast<GK*>(base + L.list1);
    GK* list2 = reinterpret_cast<GK*>(base + L.list2);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto grid_for = [&](unsigned long long items, int per_cta) {
        unsigned long long want = (items + per_cta - 1) / per_cta;
        const unsigned long long cap = (unsigned long long)sms * 8;
        return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    };
    RFI_CUDA_TRY(cudaMemsetAsync(w, 0, sizeof(GsWork), st));
    gs_pass_a_kernel<DT><<<grid_for((unsigned long long)n / 4 + 1, kGsNT * 4), kGsNT, 0, st>>>(data, flags, n, keys, w);
    gs_after_a_kernel<<<1, 1, 0, st>>>(n, w);
    const dim3 sgrid(kGsSample / 256, 2);
    // compaction levels: every level keeps ~5 % of its source (2 x 4.5 sigma of a rank in a sample of 8192);
    // with a margin of 1.5 the last list must fit one CTA's resolve
    int levels = 1;
    for (double len = (double)n * 0.075; len > (double)kGsFinal && levels < kGsMaxLevels; len *= 0.075) ++levels;
    for (int phase = 0; phase < 2; ++phase) {   // 0: median, 1: MAD (keys = |x - median|)
        GK* src = nullptr;
        unsigned long long src_cap = 0;
        for (int lv = 0; lv < levels; ++lv) {
            GK* dst = (lv & 1) ? list2 : list1;
            const unsigned long long dst_cap = (lv & 1) ? L.cap2 : L.cap1;
            if (lv == 0) {
                gs_sample_kernel<<<sgrid, 256, 0, st>>>(keys, flags, n, nullptr, 0, 0, w, sample);
                gs_bracket_kernel<<<4, 512, 0, st>>>(sample, w);
                if (phase == 0)
                    gs_compact_kernel<true, true><<<grid_for((unsigned long long)n, kGsChunk), kGsNT, 0, st>>>(keys, flags, n, nullptr, 0, dst, dst_cap, dst_cap, w);
                else
                    gs_compact_kernel<true, false><<<grid_for((unsigned long long)n, kGsChunk), kGsNT, 0, st>>>(keys, flags, n, nullptr, 0, dst, dst_cap, dst_cap, w);
            } else {
                gs_sample_kernel<<<sgrid, 256, 0, st>>>(nullptr, nullptr, 0, src, src_cap, lv, w, sample);
                gs_bracket_kernel<<<4, 512, 0, st>>>(sample, w);
                gs_compact_kernel<false, false><<<dim3(grid_for(src_cap, kGsChunk), 2), kGsNT, 0, st>>>(nullptr, nullptr, 0, src, src_cap, dst, dst_cap, dst_cap, w);
            }
            gs_advance_kernel<<<1, 1, 0, st>>>(w);
            src = dst; src_cap = dst_cap;
        }
        gs_resolve_kernel<<<2, 512, 0, st>>>(src, src_cap, w);
        if (phase == 0) gs_after_median_kernel<<<1, 1, 0, st>>>(w);
    }
    gs_finish_kernel<<<1, 1, 0, st>>>(w, out);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

// ---- baseline-sharded cube: the same statistics over the UNION of the ranks' shards --------------
// Every rank runs pass A on its shard (keys in its scratch), the moments are summed over the ranks by the
// caller (torch.distributed), and the order statistics are found by an MSB-first radix select whose
// per-pass counts -- 15 trial keys per set, one pass over the 4 B / px scratch -- are all-reduced: 8
// passes + one "next key" pass per statistic, each a few hundred bytes on the wire.
__global__ void gs_export_kernel(const GsWork* __restrict__ w, long long n, double* __restrict__ state) {
    state[0] = (double)n; state[1] = (double)w->n_flagged;
    state[2] = (double)w->n_nan[0]; state[3] = (double)w->n_nan[1];
    state[4] = w->sum[0]; state[5] = w->sum[1];
    state[6] = (double)w->maxkey[0]; state[7] = (double)w->maxkey[1];
}

template <int MODE>   // 0: counts below 15 trial keys per set; 1: count(key <= prefix), min(key > prefix)
__global__ void __launch_bounds__(kGsNT)
gs_shard_count_kernel(const GK* __restrict__ keys, const uint8_t* __restrict__ flags, long long n, int dev,
                      float centre0, float centre1, float mean0, float mean1, GK prefix0, GK prefix1, int shift,
                      int moments, unsigned long long* __restrict__ counts, double* __restrict__ sumsq) {
    const float centre[2] = {centre0, centre1}, mean[2] = {mean0, mean1};
    const GK prefix[2] = {prefix0, prefix1};
    uint32_t c[2][15];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int t = 0; t < 15; ++t) c[s][t] = 0;
    // (32-bit per-thread counters: a thread sees n / (grid * 256) keys, far below 2^32 for any cube that fits HBM)
    uint32_t cle[2] = {0, 0};
    GK nxt[2] = {kGExcl, kGExcl};
    double q[2] = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * kGsNT + threadIdx.x; i < n; i += (long long)gridDim.x * kGsNT) {
        const GK k = keys[i];
        const bool fl = flags && flags[i] != 0;
        if (moments) {
            const float x = from_key<float>(k);
            const float da = x - mean[0], dc = x - mean[1];
            q[0] += (double)(da * da);
            q[1] += fl ? 0.0 : (double)(dc * dc);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const GK x = gs_key0(k, fl, s, dev, centre[s]);
            if (MODE == 0) {
#pragma unroll
                for (int t = 0; t < 15; ++t) c[s][t] += (x < (prefix[s] | ((GK)(t + 1) << shift))) ? 1u : 0u;
            } else {
                cle[s] += (x <= prefix[s]) ? 1u : 0u;
                const GK y = x > prefix[s] ? x : kGExcl;
                nxt[s] = y < nxt[s] ? y : nxt[s];
            }
        }
    }
    if (MODE == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int t = 0; t < 15; ++t) {
                unsigned long long v = c[s][t];
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[s * 16 + t], v);
            }
    } else {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unsigned long long v = cle[s];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const GK m = warp_min(nxt[s]);
            if ((threadIdx.x & 31) == 0) {
                if (v) atomicAdd(&counts[s * 16 + 0], v);
                atomicMin(&counts[s * 16 + 1], (unsigned long long)m);
            }
        }
    }
    if (moments) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            double v = q[s];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sumsq[s], v);
        }
    }
}

}  // namespace rfi

using namespace rfi;

extern "C" int rfi_statistics_shard_begin(const void* data, int dtype, const uint8_t* flags, int64_t n,
                                          void* workspace, double* state, void* stream) {
    if (dtype != RFI_F32 && dtype != RFI_C64) { set_error("sharded statistics: float32 / complex64 data (dtype %d)", dtype); return RFI_E_UNSUPPORTED; }
    if (n < 0 || !workspace || !state || (n > 0 && !data)) { set_error("bad arguments to rfi_statistics_shard_begin"); return RFI_E_INVALID; }
    if ((reinterpret_cast<uintptr_t>(data) & 15) || (flags && (reinterpret_cast<uintptr_t>(flags) & 3))) {
        set_error("rfi_statistics_shard_begin: data must be 16-byte aligned, flags 4-byte aligned"); return RFI_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = gs_layout(n);
    char* base = static_cast<char*>(workspace);
    GsWork* w = reinterpret_cast<GsWork*>(base + L.work);
    GK* keys = reinterpret_cast<GK*>(base + L.keys);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    RFI_CUDA_TRY(cudaMemsetAsync(w, 0, sizeof(GsWork), st));
    if (n > 0) {
        unsigned long long want = ((unsigned long long)n / 4 + kGsNT * 4) / (kGsNT * 4);
        const unsigned long long cap = (unsigned long long)sms * 8;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
        if (dtype == RFI_F32) gs_pass_a_kernel<RFI_F32><<<grid, kGsNT, 0, st>>>(data, flags, n, keys, w);
        else gs_pass_a_kernel<RFI_C64><<<grid, kGsNT, 0, st>>>(data, flags, n, keys, w);
    }
    gs_export_kernel<<<1, 1, 0, st>>>(w, n, state);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

extern "C" int rfi_statistics_shard_count(const uint8_t* flags, int64_t n, void* workspace, int mode, int dev_mode,
                                          const float* centre, const float* mean, const uint32_t* prefix, int shift,
                                          unsigned long long* counts, double* sumsq, void* stream) {
    if (n < 0 || !workspace || !counts || !centre || !mean || !prefix || (mode != 0 && mode != 1) || shift < 0 || shift > 28) {
        set_error("bad arguments to rfi_statistics_shard_count"); return RFI_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = gs_layout(n);
    const GK* keys = reinterpret_cast<const GK*>(static_cast<char*>(workspace) + L.keys);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    RFI_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 32, st));
    if (mode == 1) {
        const unsigned long long ones = (unsigned long long)kGExcl;   // "no key above" (stays positive as int64 for the MIN all-reduce)
        RFI_CUDA_TRY(cudaMemcpyAsync(counts + 1, &ones, sizeof(ones), cudaMemcpyHostToDevice, st));
        RFI_CUDA_TRY(cudaMemcpyAsync(counts + 17, &ones, sizeof(ones), cudaMemcpyHostToDevice, st));
    }
    if (sumsq) RFI_CUDA_TRY(cudaMemsetAsync(sumsq, 0, sizeof(double) * 2, st));
    if (n > 0) {
        unsigned long long want = ((unsigned long long)n + kGsNT * 16 - 1) / (kGsNT * 16);
        const unsigned long long cap = (unsigned long long)sms * 8;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
        if (mode == 0)
            gs_shard_count_kernel<0><<<grid, kGsNT, 0, st>>>(keys, flags, n, dev_mode, centre[0], centre[1], mean[0], mean[1],
                                                             prefix[0], prefix[1], shift, sumsq != nullptr, counts, sumsq);
        else
            gs_shard_count_kernel<1><<<grid, kGsNT, 0, st>>>(keys, flags, n, dev_mode, centre[0], centre[1], mean[0], mean[1],
                                                             prefix[0], prefix[1], shift, 0, counts, sumsq);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

extern "C" size_t rfi_statistics2_workspace_bytes(int dtype, int64_t n) {
    if (dtype != RFI_F32 && dtype != RFI_C64) return 2 * rfi_statistics_workspace_bytes();
    return gs_layout(n < 0 ? 0 : n).total;
}

extern "C" int rfi_statistics2(const void* data, int dtype, const uint8_t* flags, int64_t n, rfi_stats_t* out,
                               void* workspace, void* stream) {
    if (n < 0 || !out || !workspace) { set_error("bad arguments to rfi_statistics2"); return RFI_E_INVALID; }
    if (n > 0 && !data) { set_error("data is NULL"); return RFI_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RFI_F32 || dtype == RFI_C64) {
        if ((reinterpret_cast<uintptr_t>(data) & 15) || (flags && (reinterpret_cast<uintptr_t>(flags) & 3))) {
            set_error("rfi_statistics2: data must be 16-byte aligned, flags 4-byte aligned");
            return RFI_E_INVALID;
        }
        return dtype == RFI_F32 ? gs_run<RFI_F32>(data, flags, n, out, workspace, st)
                                : gs_run<RFI_C64>(data, flags, n, out, workspace, st);
    }
    if (dtype != RFI_F64 && dtype != RFI_C128) { set_error("bad dtype %d", dtype); return RFI_E_INVALID; }
    // float64 arithmetic: the radix passes, once per set
    char* ws = static_cast<char*>(workspace);
    int rc = rfi_statistics(data, dtype, nullptr, n, out, ws, stream);
    if (rc) return rc;
    return rfi_statistics(data, dtype, flags, n, out + 1, ws + rfi_statistics_workspace_bytes(), stream);
}
