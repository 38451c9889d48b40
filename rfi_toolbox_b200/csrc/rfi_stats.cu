// rfi_stats.cu -- compute_statistics / compute_ffi reductions on sm_100a.
//
// Replaces rfi_toolbox/evaluation/statistics.py:10-56: |z| (fused into the load), the boolean
// gather data[~flags], mean, population std, median and MAD (= median(|x - median|)).
//
// One call produces the statistics of the unflagged samples (or of all samples when flags is
// NULL).  Everything is enqueued on the caller's stream with no host synchronisation:
//   moments pass 1   count, flagged count, NaN count, sum (fp64 accumulators)
//   moments pass 2   sum of squared deviations from the T-rounded mean
//   select           exact order statistics by MSB-first radix select, 4 bits per pass:
//                    each pass counts, for 15 trial keys, how many samples lie below
//                    (register counters, no shared-memory atomics), a one-thread kernel then
//                    fixes the next 4 bits.  8 passes for float32 keys, 16 for float64.
// The median is the mean of the two middle order statistics in T, as NumPy forms it.
#include "rfi_common.cuh"

namespace rfi {

struct StatsWork {
    // moments
    unsigned long long n_clean, n_flagged, n_nan;
    double sum, sumsq;
    unsigned long long maxkey;  // ordered key (of the double-widened value) of the largest sample; NaN = all ones
    double mean_t;  // mean rounded to T
    // select state
    unsigned long long prefix, k1, k2, n_sel;
    unsigned long long cnt[16];
    unsigned long long cle, nxt;  // count(key <= prefix), min(key > prefix)
    double centre;                // median (T-rounded) used by the MAD select
    int shift, pad;
    // result
    rfi_stats_t out;
};

template <int DT> struct SIn;
template <> struct SIn<RFI_F32>  { using T = float;  };
template <> struct SIn<RFI_F64>  { using T = double; };
template <> struct SIn<RFI_C64>  { using T = float;  };
template <> struct SIn<RFI_C128> { using T = double; };

template <int DT>
RFI_DEVINL typename SIn<DT>::T load_mag(const void* base, long long i) {
    using T = typename SIn<DT>::T;
    if constexpr (DT == RFI_F32) return __ldg(static_cast<const float*>(base) + i);
    else if constexpr (DT == RFI_F64) return __ldg(static_cast<const double*>(base) + i);
    else if constexpr (DT == RFI_C64) { float2 z = __ldg(static_cast<const float2*>(base) + i); return cabs_np<T>(z.x, z.y); }
    else { double2 z = __ldg(static_cast<const double2*>(base) + i); return cabs_np<T>(z.x, z.y); }
}

constexpr int kStatThreads = 256;

RFI_DEVINL double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
RFI_DEVINL unsigned long long warp_sum_u(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int DT>
__global__ void __launch_bounds__(kStatThreads)
moments1_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long n, StatsWork* w) {
    using T = typename SIn<DT>::T;
    unsigned long long nc = 0, nf = 0, nn = 0, mk = 0;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kStatThreads) {
        const bool fl = flags ? (__ldg(flags + i) != 0) : false;
        nf += fl;
        if (!fl) {
            T x = load_mag<DT>(data, i);
            nc += 1; nn += is_nan(x) ? 1 : 0;
            s += (double)x;
            const unsigned long long k = to_key<double>((double)x);
            mk = k > mk ? k : mk;
        }
    }
    nc = warp_sum_u(nc); nf = warp_sum_u(nf); nn = warp_sum_u(nn); s = warp_sum_d(s);
    mk = warp_max(mk);
    if ((threadIdx.x & 31) == 0) {
        if (mk) atomicMax(&w->maxkey, mk);
        if (nc) atomicAdd(&w->n_clean, nc);
        if (nf) atomicAdd(&w->n_flagged, nf);
        if (nn) atomicAdd(&w->n_nan, nn);
        atomicAdd(&w->sum, s);
    }
}

template <typename T>
__global__ void mean_kernel(StatsWork* w) {
    w->mean_t = w->n_clean ? (double)(T)(w->sum / (double)w->n_clean) : 0.0;
}

template <int DT>
__global__ void __launch_bounds__(kStatThreads)
moments2_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long n, StatsWork* w) {
    using T = typename SIn<DT>::T;
    const T mean = (T)w->mean_t;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kStatThreads) {
        if (flags && __ldg(flags + i) != 0) continue;
        T d = load_mag<DT>(data, i) - mean;  // np.std: abs(x - mean)**2 in T, then summed
        s += (double)(d * d);
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) atomicAdd(&w->sumsq, s);
}

// key of sample i for the current select: the magnitude itself (MODE 0) or its absolute
// deviation from the centre (MODE 1).  Flagged samples and NaNs carry the all-ones key.
template <int DT, int MODE>
RFI_DEVINL typename Scalar<typename SIn<DT>::T>::key_t sel_key(const void* data, const uint8_t* flags,
                                                               long long i, typename SIn<DT>::T centre) {
    using T = typename SIn<DT>::T;
    using K = typename Scalar<T>::key_t;
    if (flags && __ldg(flags + i) != 0) return ~K(0);
    T x = load_mag<DT>(data, i);
    if (MODE == 1) x = fabs_(x - centre);
    return to_key<T>(x);
}

template <int DT, int MODE>
__global__ void __launch_bounds__(kStatThreads)
select_count_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long n, StatsWork* w) {
    using T = typename SIn<DT>::T;
    using K = typename Scalar<T>::key_t;
    const K prefix = (K)w->prefix;
    const int shift = w->shift;
    const T centre = (T)w->centre;
    uint32_t c[15];
#pragma unroll
    for (int t = 0; t < 15; ++t) c[t] = 0;
    unsigned long long big[15];
#pragma unroll
    for (int t = 0; t < 15; ++t) big[t] = 0;
    unsigned it = 0;
    for (long long i = (long long)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kStatThreads) {
        const K key = sel_key<DT, MODE>(data, flags, i, centre);
#pragma unroll
        for (int t = 0; t < 15; ++t) c[t] += (key < (prefix | ((K)(t + 1) << shift))) ? 1u : 0u;
        if ((++it & 0xffffu) == 0) {
#pragma unroll
            for (int t = 0; t < 15; ++t) { big[t] += c[t]; c[t] = 0; }
        }
    }
    __shared__ unsigned long long sh[15];
    if (threadIdx.x < 15) sh[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 15; ++t) {
        unsigned long long v = warp_sum_u(big[t] + c[t]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[t], v);
    }
    __syncthreads();
    if (threadIdx.x < 15 && sh[threadIdx.x]) atomicAdd(&w->cnt[threadIdx.x], sh[threadIdx.x]);
}

// fix the next 4 bits: digit = number of trials whose below-count is <= k1
__global__ void select_decide_kernel(StatsWork* w) {
    int d = 0;
    for (int t = 0; t < 15; ++t) d += (w->cnt[t] <= w->k1) ? 1 : 0;
    w->prefix |= (unsigned long long)d << w->shift;
    w->shift -= 4;
    for (int t = 0; t < 16; ++t) w->cnt[t] = 0;
}

template <int DT, int MODE>
__global__ void __launch_bounds__(kStatThreads)
select_next_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long n, StatsWork* w) {
    using T = typename SIn<DT>::T;
    using K = typename Scalar<T>::key_t;
    const K prefix = (K)w->prefix;
    const T centre = (T)w->centre;
    unsigned long long cle = 0;
    K nxt = ~K(0);
    for (long long i = (long long)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kStatThreads) {
        const K key = sel_key<DT, MODE>(data, flags, i, centre);
        cle += (key <= prefix) ? 1 : 0;
        const K y = key > prefix ? key : ~K(0);
        nxt = y < nxt ? y : nxt;
    }
    cle = warp_sum_u(cle);
    nxt = warp_min(nxt);
    if ((threadIdx.x & 31) == 0) {
        if (cle) atomicAdd(&w->cle, cle);
        atomicMin(&w->nxt, (unsigned long long)nxt);
    }
}

// median of the selected set from (prefix, cle, nxt); MODE 0 stores it as the centre of the
// following MAD select, MODE 1 stores the MAD.
template <typename T, int MODE>
__global__ void select_finish_kernel(StatsWork* w) {
    using K = typename Scalar<T>::key_t;
    T med = Scalar<T>::nan();
    if (w->n_sel > 0) {
        const T a = from_key<T>((K)w->prefix);
        T b = a;
        if (w->k2 != w->k1) b = (w->k2 < w->cle) ? a : from_key<T>((K)w->nxt);
        med = (w->k1 == w->k2) ? a : (a + b) * T(0.5);
    }
    if (MODE == 0) { w->centre = (double)med; w->out.median = (double)med; }
    else w->out.mad = (double)med;
}

__global__ void select_begin_kernel(StatsWork* w, int key_bits) {
    const unsigned long long n = w->n_clean - w->n_nan;  // NaNs never rank (they poison the result on the host side)
    w->n_sel = n;
    w->k1 = n ? (n - 1) >> 1 : 0;
    w->k2 = n >> 1;
    w->prefix = 0;
    w->shift = key_bits - 4;
    for (int t = 0; t < 16; ++t) w->cnt[t] = 0;
    w->cle = 0;
    w->nxt = ~0ull;
}

template <typename T>
__global__ void stats_finish_kernel(StatsWork* w) {
    rfi_stats_t& o = w->out;
    o.count = (long long)w->n_clean;
    o.n_flagged = (long long)w->n_flagged;
    o.n_nan = (long long)w->n_nan;
    if (w->n_clean == 0) {
        o.mean = o.std = o.median = o.mad = o.max = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    o.max = from_key<double>(w->maxkey);
    o.mean = w->mean_t;
    const T var = (T)(w->sumsq / (double)w->n_clean);
    o.std = (double)Scalar<T>::sqrt_rn(var);
    if (w->n_nan) o.median = o.mad = __longlong_as_double(0x7ff8000000000000LL);  // np.median propagates NaN
}

template <int DT>
static int run_stats(const void* data, const uint8_t* flags, long long n, StatsWork* w, cudaStream_t st) {
    using T = typename SIn<DT>::T;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (n + kStatThreads * 4 - 1) / (kStatThreads * 4);
    long long cap = (long long)sms * 8;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    constexpr int kBits = sizeof(T) * 8;

    RFI_CUDA_TRY(cudaMemsetAsync(w, 0, sizeof(StatsWork), st));
    moments1_kernel<DT><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
    mean_kernel<T><<<1, 1, 0, st>>>(w);
    moments2_kernel<DT><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
    // median
    select_begin_kernel<<<1, 1, 0, st>>>(w, kBits);
    for (int p = 0; p < kBits / 4; ++p) {
        select_count_kernel<DT, 0><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
        select_decide_kernel<<<1, 1, 0, st>>>(w);
    }
    select_next_kernel<DT, 0><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
    select_finish_kernel<T, 0><<<1, 1, 0, st>>>(w);
    // MAD
    select_begin_kernel<<<1, 1, 0, st>>>(w, kBits);
    for (int p = 0; p < kBits / 4; ++p) {
        select_count_kernel<DT, 1><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
        select_decide_kernel<<<1, 1, 0, st>>>(w);
    }
    select_next_kernel<DT, 1><<<grid, kStatThreads, 0, st>>>(data, flags, n, w);
    select_finish_kernel<T, 1><<<1, 1, 0, st>>>(w);
    stats_finish_kernel<T><<<1, 1, 0, st>>>(w);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

// ------------------------------------------------------------------------------------------
// Per-pair sweep (BASELINE config 4): statistics of every consecutive segment of `seg` samples
// (one 128 x 128 patch pair per CTA), before and after flagging, in ONE launch.  The segment
// lives in registers as order-preserving keys (32 per thread); medians / MADs by the
// register-resident radix select of rfi_common.cuh; moments accumulate in float64.
constexpr int kSegNT = 512, kSegE = 32;  // <= 16384 samples per segment

template <typename T, int NT>
RFI_DEVINL void seg_moments(const T (&x)[kSegE], const bool (&use)[kSegE], double* red, rfi_stats_t& o,
                            T& mean_t, uint32_t& n_use, uint32_t& n_nan) {
    // red: shared scratch of 5 doubles (zeroed by the caller before a barrier)
    double s = 0.0;
    uint32_t n = 0, nn = 0;
    unsigned long long mk = 0;
#pragma unroll
    for (int e = 0; e < kSegE; ++e) {
        if (use[e]) {
            s += (double)x[e]; n += 1; nn += is_nan(x[e]) ? 1u : 0u;
            const unsigned long long k = to_key<double>((double)x[e]);
            mk = k > mk ? k : mk;
        }
    }
    mk = warp_max(mk);
    if ((threadIdx.x & 31) == 0 && mk) atomicMax(reinterpret_cast<unsigned long long*>(&red[4]), mk);
    s = warp_sum_d(s);
    n = __reduce_add_sync(0xffffffffu, n);
    nn = __reduce_add_sync(0xffffffffu, nn);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&red[0], s);
        atomicAdd(&red[1], (double)n);
        atomicAdd(&red[2], (double)nn);
    }
    __syncthreads();
    const double sum = red[0];
    n_use = (uint32_t)red[1];
    n_nan = (uint32_t)red[2];
    mean_t = n_use ? (T)(sum / (double)n_use) : T(0);
    double q = 0.0;
#pragma unroll
    for (int e = 0; e < kSegE; ++e) {
        if (use[e]) { const T d = x[e] - mean_t; q += (double)(d * d); }
    }
    q = warp_sum_d(q);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[3], q);
    __syncthreads();
    const double sumsq = red[3];
    o.count = (long long)n_use;
    o.n_nan = (long long)n_nan;
    if (n_use == 0) {
        o.mean = o.std = o.median = o.mad = o.max = __longlong_as_double(0x7ff8000000000000LL);
    } else {
        o.max = from_key<double>(*reinterpret_cast<unsigned long long*>(&red[4]));
        o.mean = (double)mean_t;
        o.std = (double)Scalar<T>::sqrt_rn((T)(sumsq / (double)n_use));
    }
    __syncthreads();
}

template <int DT>
__global__ void __launch_bounds__(kSegNT, 1)
stats_segmented_kernel(const void* __restrict__ data, const uint8_t* __restrict__ flags, long long seg,
                       rfi_stats_t* __restrict__ out) {
    using T = typename SIn<DT>::T;
    using K = typename Scalar<T>::key_t;
    constexpr int NT = kSegNT, E = kSegE;
    constexpr K kExcl = ~K(0);
    __shared__ BlockScratch<NT> scr;
    __shared__ RoundCounter rc;
    __shared__ SelectScratch<K> sel;
    __shared__ double red[5];  // sum, count, NaNs, squared deviations, max key (bits)
    int parity = 0, round = 0;
    round_init(rc);
    const long long base = (long long)blockIdx.x * seg;
    T x[E];
    bool in_seg[E], clean[E];
    uint32_t nflag = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const long long i = (long long)e * NT + threadIdx.x;
        in_seg[e] = i < seg;
        x[e] = in_seg[e] ? load_mag<DT>(data, base + i) : T(0);
        const bool fl = in_seg[e] && flags && __ldg(flags + base + i) != 0;
        nflag += fl ? 1u : 0u;
        clean[e] = in_seg[e] && !fl;
    }
    for (int which = 0; which < 2; ++which) {  // 0: all samples, 1: unflagged samples
        const bool (&use)[E] = which == 0 ? in_seg : clean;
        rfi_stats_t o;
        o.n_flagged = 0;
        if (threadIdx.x < 5) red[threadIdx.x] = 0.0;
        __syncthreads();
        T mean_t;
        uint32_t n_use, n_nan;
        seg_moments<T, NT>(x, use, red, o, mean_t, n_use, n_nan);
        if (n_use > 0) {
            if (n_nan > 0) {  // np.median propagates NaN
                o.median = o.mad = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                K key[E];
#pragma unroll
                for (int e = 0; e < E; ++e) key[e] = use[e] ? to_key<T>(x[e]) : kExcl;
                const T med = block_median<T, NT, E>(key, n_use, scr, parity, rc, round, sel);
#pragma unroll
                for (int e = 0; e < E; ++e) key[e] = use[e] ? to_key<T>(fabs_(x[e] - med)) : kExcl;
                const T mad = block_median<T, NT, E>(key, n_use, scr, parity, rc, round, sel);
                o.median = (double)med;
                o.mad = (double)mad;
            }
        }
        if (which == 1) {
            const uint32_t nf = round_sum(nflag, rc, round);
            o.n_flagged = (long long)nf;
        }
        if (threadIdx.x == 0) out[(long long)blockIdx.x * 2 + which] = o;
        __syncthreads();
    }
}

}  // namespace rfi

extern "C" size_t rfi_statistics_workspace_bytes(void) { return sizeof(rfi::StatsWork); }

extern "C" int rfi_statistics(const void* data, int dtype, const uint8_t* flags, int64_t n,
                              rfi_stats_t* out, void* workspace, void* stream) {
    using namespace rfi;
    if (n < 0 || !out || !workspace) { set_error("bad arguments to rfi_statistics"); return RFI_E_INVALID; }
    if (n > 0 && !data) { set_error("data is NULL"); return RFI_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    StatsWork* w = static_cast<StatsWork*>(workspace);
    int rc;
    switch (dtype) {
        case RFI_F32:  rc = run_stats<RFI_F32>(data, flags, n, w, st); break;
        case RFI_F64:  rc = run_stats<RFI_F64>(data, flags, n, w, st); break;
        case RFI_C64:  rc = run_stats<RFI_C64>(data, flags, n, w, st); break;
        case RFI_C128: rc = run_stats<RFI_C128>(data, flags, n, w, st); break;
        default: set_error("bad dtype %d", dtype); return RFI_E_INVALID;
    }
    if (rc) return rc;
    RFI_CUDA_TRY(cudaMemcpyAsync(out, &w->out, sizeof(rfi_stats_t), cudaMemcpyDeviceToDevice, st));
    return RFI_OK;
}

extern "C" int rfi_statistics_segmented(const void* data, int dtype, const uint8_t* flags, int64_t n_seg,
                                        int64_t seg, rfi_stats_t* out, void* stream) {
    using namespace rfi;
    if (n_seg < 0 || seg <= 0 || (n_seg > 0 && (!data || !out))) { set_error("bad arguments to rfi_statistics_segmented"); return RFI_E_INVALID; }
    // float32 arithmetic: the sampled-bracket kernel of rfi_pairs.cu (2 CTAs / SM, ~6x the rate of the
    // register-resident radix selects below, which stay for float64 / complex128)
    if ((dtype == RFI_F32 || dtype == RFI_C64) && seg <= (int64_t)kSegNT * kSegE)
        return rfi_pair_sweep(data, dtype, flags, nullptr, n_seg, seg, out, nullptr, stream);
    if (seg > (int64_t)kSegNT * kSegE) {
        set_error("segment of %lld samples: the per-pair kernel holds at most %d (use rfi_statistics per segment)",
                  (long long)seg, kSegNT * kSegE);
        return RFI_E_UNSUPPORTED;
    }
    if (n_seg == 0) return RFI_OK;
    if (n_seg > 0x7fffffffLL) { set_error("too many segments for one launch"); return RFI_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_seg;
    switch (dtype) {
        case RFI_F32:  stats_segmented_kernel<RFI_F32><<<grid, kSegNT, 0, st>>>(data, flags, seg, out); break;
        case RFI_F64:  stats_segmented_kernel<RFI_F64><<<grid, kSegNT, 0, st>>>(data, flags, seg, out); break;
        case RFI_C64:  stats_segmented_kernel<RFI_C64><<<grid, kSegNT, 0, st>>>(data, flags, seg, out); break;
        case RFI_C128: stats_segmented_kernel<RFI_C128><<<grid, kSegNT, 0, st>>>(data, flags, seg, out); break;
        default: set_error("bad dtype %d", dtype); return RFI_E_INVALID;
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
