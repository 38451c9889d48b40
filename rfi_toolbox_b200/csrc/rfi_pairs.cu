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
#include "rfi_select_keys.cuh"

namespace rfi {

constexpr int kPairNT = kSelNT, kPairE = 32, kPairG = kPairE / 4, kPairSeg = kPairNT * kPairE;
constexpr PK kPExcl = kSelExcl;

struct PairShared {
    SelShared sel;
    double pd[4][16];     // per-warp partials
    uint32_t pu[8][16];
    double rd[4];         // block totals
    uint32_t ru[8];
    uint32_t ext[4];      // the two maxima and the two minima of the load pass
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

struct PairKeys {   // all eight groups in shared memory, [g][thread][4]
    const PK* skeys;
    RFI_DEVINL uint4 load(int g) const { return *reinterpret_cast<const uint4*>(skeys + ((size_t)g * kPairNT + threadIdx.x) * 4); }
    RFI_DEVINL PK load1(int g, int i) const { return skeys[((size_t)g * kPairNT + threadIdx.x) * 4 + i]; }
};

// median and MAD of the nv valid keys (nv >= 1, no NaN among them)
__device__ __noinline__ void pair_median_mad(const PK* __restrict__ skeys, PK* cand, PK* samp, PairShared& sh,
                                             uint32_t nv, bool any_inf, float& med, float& mad) {
    const PairKeys ka{skeys};
    PK a = 0, b = 0;
    int sv = 0;
    sel_median(ka, cand, samp, sh.sel, nv, !any_inf, a, b, sv);
    med = median_of_pair<float>(from_key<float>(a), from_key<float>(b), nv);
    if (is_inf(med) || is_nan(med)) {  // some |x - med| is inf - inf = NaN: np.median propagates it
        mad = Scalar<float>::nan();
        return;
    }
    sel_mad(ka, cand, samp, sh.sel, nv, sv, med, a, b);
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
    if (tid == 0) { sh.ext[0] = 0; sh.ext[1] = 0; sh.ext[2] = kPExcl; sh.ext[3] = kPExcl; }
    __syncthreads();
    if ((tid & 31) == 0) {
        atomicMax(&sh.ext[0], mx_all); atomicMax(&sh.ext[1], mx_cln);
        atomicMin(&sh.ext[2], mn_all); atomicMin(&sh.ext[3], mn_cln);
    }
    double d2[4] = {s_all, s_cln, ss_all, ss_cln};
    uint32_t u6[6] = {nflag, tp, fp, fn, nan_all, nan_cln};
    pair_totals<4, 6>(d2, u6, sh);   // (its barriers also publish the extremes)
    mx_all = sh.ext[0]; mx_cln = sh.ext[1];
    mn_all = sh.ext[2]; mn_cln = sh.ext[3];
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
