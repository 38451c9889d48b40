// rfi_metrics.cu -- confusion counts for evaluate_segmentation on sm_100a.
//
// Replaces the 12 astype(bool) copies and 17 boolean reduction passes of
// rfi_toolbox/evaluation/metrics.py:25-172 with one streaming pass: 128-bit loads of both
// masks, per-thread popcounts, warp-shuffle + block reduction, one 64-bit atomic per CTA.
// The five ratios are formed on the host in float64 exactly as the reference does.
#include <string.h>

#include <stdlib.h>

#include "rfi_common.cuh"

namespace rfi {

// "truthiness" of one element, as ndarray.astype(bool): non-zero (NaN included) is True.
template <int ELEM, bool FLT> struct Truth;
template <> struct Truth<1, false> { using V = uint8_t;  RFI_DEVINL static bool nz(V x) { return x != 0; } };
template <> struct Truth<2, false> { using V = uint16_t; RFI_DEVINL static bool nz(V x) { return x != 0; } };
template <> struct Truth<4, false> { using V = uint32_t; RFI_DEVINL static bool nz(V x) { return x != 0; } };
template <> struct Truth<8, false> { using V = unsigned long long; RFI_DEVINL static bool nz(V x) { return x != 0; } };
// IEEE: +-0 is False, everything else (NaN too) True -> test the magnitude bits
template <> struct Truth<2, true> { using V = uint16_t; RFI_DEVINL static bool nz(V x) { return (x & 0x7fffu) != 0; } };
template <> struct Truth<4, true> { using V = uint32_t; RFI_DEVINL static bool nz(V x) { return (x & 0x7fffffffu) != 0; } };
template <> struct Truth<8, true> { using V = unsigned long long; RFI_DEVINL static bool nz(V x) { return (x & 0x7fffffffffffffffull) != 0; } };

// bit mask (one bit per element) of the non-zero elements among the 16 bytes at p;
// returns the number of elements covered (16 / ELEM).
template <int ELEM, bool FLT>
RFI_DEVINL uint32_t nz_mask16(const uint4& q) {
    if constexpr (ELEM == 1) {
        // high bit of every byte set iff the byte is non-zero, then gather those bits
        auto hb = [](uint32_t v) { return (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u; };
        auto pack = [](uint32_t h) {  // bits 7,15,23,31 -> 0..3
            return ((h >> 7) & 1u) | ((h >> 14) & 2u) | ((h >> 21) & 4u) | ((h >> 28) & 8u);
        };
        return pack(hb(q.x)) | (pack(hb(q.y)) << 4) | (pack(hb(q.z)) << 8) | (pack(hb(q.w)) << 12);
    } else if constexpr (ELEM == 2) {
        using TR = Truth<2, FLT>;
        uint32_t w[4] = {q.x, q.y, q.z, q.w}, m = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m |= (TR::nz((uint16_t)(w[i] & 0xffffu)) ? 1u : 0u) << (2 * i);
            m |= (TR::nz((uint16_t)(w[i] >> 16)) ? 1u : 0u) << (2 * i + 1);
        }
        return m;
    } else if constexpr (ELEM == 4) {
        using TR = Truth<4, FLT>;
        return (TR::nz(q.x) ? 1u : 0u) | (TR::nz(q.y) ? 2u : 0u) | (TR::nz(q.z) ? 4u : 0u) | (TR::nz(q.w) ? 8u : 0u);
    } else {
        using TR = Truth<8, FLT>;
        unsigned long long a = ((unsigned long long)q.y << 32) | q.x, b = ((unsigned long long)q.w << 32) | q.z;
        return (TR::nz(a) ? 1u : 0u) | (TR::nz(b) ? 2u : 0u);
    }
}

template <int ELEM, bool FLT>
RFI_DEVINL bool nz_scalar(const void* base, long long i) {
    using TR = Truth<ELEM, FLT>;
    return TR::nz(static_cast<const typename TR::V*>(base)[i]);
}

// Both masks with the SAME element size: the common case (uint8/bool vs uint8/bool).
// `n` elements starting at (pred, truth); segment handled by the calling block(s).
template <int EP, bool FP, int ET, bool FT>
RFI_DEVINL void count_range(const void* pred, const void* truth, long long begin, long long end,
                            long long tid, long long nthreads,
                            unsigned long long& tp, unsigned long long& fp, unsigned long long& fn) {
    constexpr int VP = 16 / EP, VT = 16 / ET;          // elements per 128-bit load
    constexpr int V = VP > VT ? VP : VT;               // elements per step (coarser of the two)
    const uintptr_t ap = reinterpret_cast<uintptr_t>(pred) + (uintptr_t)begin * EP;
    const uintptr_t at = reinterpret_cast<uintptr_t>(truth) + (uintptr_t)begin * ET;
    const long long n = end - begin;
    // scalar head until BOTH operands sit on a 16-byte boundary; if no such head exists
    // (the two arrays have different phases) the whole range goes through the scalar path
    long long head = n;
    for (int h = 0; h < 16; ++h) {
        if ((((ap + (uintptr_t)h * EP) | (at + (uintptr_t)h * ET)) & 15) == 0) { head = h; break; }
    }
    if (head > n) head = n;
    unsigned tpc = 0, fpc = 0, fnc = 0;
    for (long long i = tid; i < head; i += nthreads) {
        bool p = nz_scalar<EP, FP>(pred, begin + i), t = nz_scalar<ET, FT>(truth, begin + i);
        tpc += (p && t); fpc += (p && !t); fnc += (!p && t);
    }
    const long long body = (n - head) / V;
    const uint4* vp = reinterpret_cast<const uint4*>(ap + (uintptr_t)head * EP);
    const uint4* vt = reinterpret_cast<const uint4*>(at + (uintptr_t)head * ET);
    constexpr int LP = V / VP, LT = V / VT;  // 128-bit loads per step for each operand
    unsigned iter = 0;
    long long s0 = tid;
    if constexpr (EP == 1 && ET == 1) {
        // byte masks (the common case): four 128-bit loads per operand in flight per thread, and the
        // counts taken byte-parallel -- high bit of every non-zero byte, three popcounts per word
        constexpr int U = 4;
        // 0x01 in every non-zero byte; the three sums are then byte dot products (dp4a): p.t, p.1, t.1
        auto nz01 = [](uint32_t v) { return ((((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u) >> 7; };
        unsigned npc = 0, ntc = 0;  // predicted / true positives; FP = P - TP, FN = T - TP
        for (; s0 + (U - 1) * nthreads < body; s0 += U * nthreads) {
            uint4 a[U], b[U];
#pragma unroll
            for (int k = 0; k < U; ++k) { a[k] = __ldg(vp + s0 + k * nthreads); b[k] = __ldg(vt + s0 + k * nthreads); }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const uint32_t pw[4] = {a[k].x, a[k].y, a[k].z, a[k].w}, tw[4] = {b[k].x, b[k].y, b[k].z, b[k].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t p = nz01(pw[i]), t = nz01(tw[i]);
                    tpc = __dp4a(p, t, tpc); npc = __dp4a(p, 0x01010101u, npc); ntc = __dp4a(t, 0x01010101u, ntc);
                }
            }
            if ((++iter & 1023) == 0) {
                tp += tpc; fp += npc - tpc; fn += ntc - tpc;
                tpc = npc = ntc = 0;
            }
        }
        tp += tpc; fp += npc - tpc; fn += ntc - tpc;
        tpc = 0;
    }
    for (long long s = s0; s < body; s += nthreads) {
        uint32_t mp = 0, mt = 0;
#pragma unroll
        for (int k = 0; k < LP; ++k) mp |= nz_mask16<EP, FP>(__ldg(vp + s * LP + k)) << (k * VP);
#pragma unroll
        for (int k = 0; k < LT; ++k) mt |= nz_mask16<ET, FT>(__ldg(vt + s * LT + k)) << (k * VT);
        tpc += __popc(mp & mt); fpc += __popc(mp & ~mt); fnc += __popc(~mp & mt);
        if ((++iter & 1023) == 0) { tp += tpc; fp += fpc; fn += fnc; tpc = fpc = fnc = 0; }
    }
    for (long long i = head + body * V + tid; i < n; i += nthreads) {
        bool p = nz_scalar<EP, FP>(pred, begin + i), t = nz_scalar<ET, FT>(truth, begin + i);
        tpc += (p && t); fpc += (p && !t); fnc += (!p && t);
    }
    tp += tpc; fp += fpc; fn += fnc;
}

RFI_DEVINL unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NT>
RFI_DEVINL void block_sum3(unsigned long long& a, unsigned long long& b, unsigned long long& c) {
    __shared__ unsigned long long sh[3][NT / 32];
    a = warp_sum64(a); b = warp_sum64(b); c = warp_sum64(c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sh[2][warp] = c; }
    __syncthreads();
    if (warp == 0) {
        a = lane < NT / 32 ? sh[0][lane] : 0; b = lane < NT / 32 ? sh[1][lane] : 0; c = lane < NT / 32 ? sh[2][lane] : 0;
        a = warp_sum64(a); b = warp_sum64(b); c = warp_sum64(c);
    }
}

constexpr int kMetricThreads = 256;

// ---- counts + all-reduce in ONE kernel, over NVLink peer memory -----------------------------
// evaluate_segmentation over a baseline-sharded cube needs the SUM of {TP, FP, FN} over the
// ranks: a reduction followed by a collective.  Every rank owns one small
// exchange buffer (cudaMalloc + CUDA IPC, opened by all ranks of the box); the last CTA of the
// reduction stores the rank's three totals into slot [rank] of EVERY peer's buffer (plain stores
// over NVLink), publishes the call's epoch in the peer's flag word behind a system fence, waits
// for the same epoch from every peer in its OWN buffer and writes the sums -- no NCCL launch, no
// extra stream hop, one kernel on the caller's stream.  Slots and flags are double-buffered by
// epoch parity: a rank can only be one call ahead of a peer that still has to read its slots.
constexpr int kPeerMax = 16;
struct PeerBuf {
    unsigned long long partial[2][4];          // this rank's CTAs accumulate here
    unsigned int ticket[2], pad[2];
    unsigned long long slots[2][kPeerMax][4];  // [parity][source rank] = {TP, FP, FN, -}
    unsigned long long flag[2][kPeerMax];      // epoch published by the source rank
    unsigned long long error;                  // epoch of a timed-out wait (0 = none)
};
struct PeerArgs {
    PeerBuf* peer[kPeerMax];
    int world, rank;
    unsigned long long epoch;
};
constexpr long long kPeerTimeoutCycles = 60000000000LL;  // ~30 s: a hung peer must not hang the GPU

template <int EP, bool FP, int ET, bool FT>
__global__ void __launch_bounds__(kMetricThreads)
confusion_kernel(const void* __restrict__ pred, const void* __restrict__ truth, long long n,
                 unsigned long long* __restrict__ counts, PeerArgs pa) {
    unsigned long long tp = 0, fp = 0, fn = 0;
    count_range<EP, FP, ET, FT>(pred, truth, 0, n, (long long)blockIdx.x * kMetricThreads + threadIdx.x,
                                (long long)gridDim.x * kMetricThreads, tp, fp, fn);
    block_sum3<kMetricThreads>(tp, fp, fn);
    if (pa.world == 0) {
        if (threadIdx.x == 0) {
            if (tp) atomicAdd(counts + 0, tp);
            if (fp) atomicAdd(counts + 1, fp);
            if (fn) atomicAdd(counts + 2, fn);
        }
        return;
    }
    const int par = (int)(pa.epoch & 1ull);
    PeerBuf* me = pa.peer[pa.rank];
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        if (tp) atomicAdd(&me->partial[par][0], tp);
        if (fp) atomicAdd(&me->partial[par][1], fp);
        if (fn) atomicAdd(&me->partial[par][2], fn);
        __threadfence();
        is_last = atomicAdd(&me->ticket[par], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    unsigned long long a = 0, b = 0, c = 0, missing = 0;
    if ((int)threadIdx.x < pa.world) {   // one thread per peer (world <= 16: all inside warp 0)
        const int q = threadIdx.x;
        volatile unsigned long long* mine = me->partial[par];
        const unsigned long long t0 = mine[0], t1 = mine[1], t2 = mine[2];
        PeerBuf* dst = pa.peer[q];
        volatile unsigned long long* sl = dst->slots[par][pa.rank];
        sl[0] = t0; sl[1] = t1; sl[2] = t2;
        __threadfence_system();
        *(volatile unsigned long long*)&dst->flag[par][pa.rank] = pa.epoch;
        volatile unsigned long long* f = &me->flag[par][q];
        const long long start = clock64();
        bool ok = true;
        while (*f != pa.epoch) {
            if (clock64() - start > kPeerTimeoutCycles) { ok = false; break; }
            __nanosleep(64);
        }
        __threadfence_system();
        if (ok) {
            volatile unsigned long long* r = me->slots[par][q];
            a = r[0]; b = r[1]; c = r[2];
        } else {
            missing = 1;
        }
    }
    if (threadIdx.x < 32) {
        a = warp_sum64(a); b = warp_sum64(b); c = warp_sum64(c); missing = warp_sum64(missing);
        if (threadIdx.x == 0) {
            counts[0] = a; counts[1] = b; counts[2] = c;
            counts[3] = missing;               // peers that never arrived
            if (missing) me->error = pa.epoch;
            me->partial[par][0] = me->partial[par][1] = me->partial[par][2] = 0;  // next use: epoch + 2
            me->ticket[par] = 0;
        }
    }
}

// one CTA per segment (a 128 x 128 pair is 16 KiB per mask)
template <int EP, bool FP, int ET, bool FT>
__global__ void __launch_bounds__(kMetricThreads)
confusion_segmented_kernel(const void* __restrict__ pred, const void* __restrict__ truth, long long n_seg,
                           long long seg, unsigned long long* __restrict__ counts) {
    for (long long s = blockIdx.x; s < n_seg; s += gridDim.x) {
        unsigned long long tp = 0, fp = 0, fn = 0;
        count_range<EP, FP, ET, FT>(pred, truth, s * seg, (s + 1) * seg, threadIdx.x, kMetricThreads, tp, fp, fn);
        block_sum3<kMetricThreads>(tp, fp, fn);
        if (threadIdx.x == 0) { counts[s * 3 + 0] = tp; counts[s * 3 + 1] = fp; counts[s * 3 + 2] = fn; }
        __syncthreads();
    }
}

template <int EP, bool FP, int ET, bool FT>
static int launch_confusion(const void* pred, const void* truth, long long n, long long n_seg, long long seg,
                            unsigned long long* counts, const PeerArgs& pa, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (n_seg < 0) {
        long long want = (n / (16 / (EP < ET ? EP : ET)) + kMetricThreads * 8 - 1) / (kMetricThreads * 8);
        // CTAs per SM: 8 x 256 threads fill an SM's thread slots; RFI_CONFUSION_CTAS_PER_SM (experiment knob)
        // leaves room for another kernel's CTAs when the counts run on a side stream
        static const int per_sm = getenv("RFI_CONFUSION_CTAS_PER_SM") ? atoi(getenv("RFI_CONFUSION_CTAS_PER_SM")) : 8;
        long long cap = (long long)sms * (per_sm > 0 ? per_sm : 8);
        unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
        confusion_kernel<EP, FP, ET, FT><<<grid, kMetricThreads, 0, st>>>(pred, truth, n, counts, pa);
    } else {
        long long cap = (long long)sms * 8;
        unsigned grid = (unsigned)(n_seg < cap ? n_seg : cap);
        confusion_segmented_kernel<EP, FP, ET, FT><<<grid, kMetricThreads, 0, st>>>(pred, truth, n_seg, seg, counts);
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

template <int EP, bool FP>
static int dispatch_true(const void* pred, const void* truth, int et, int ft, long long n, long long n_seg,
                         long long seg, unsigned long long* counts, const PeerArgs& pa, cudaStream_t st) {
    switch (et * 2 + (ft ? 1 : 0)) {
        case 2:  return launch_confusion<EP, FP, 1, false>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 4:  return launch_confusion<EP, FP, 2, false>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 5:  return launch_confusion<EP, FP, 2, true>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 8:  return launch_confusion<EP, FP, 4, false>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 9:  return launch_confusion<EP, FP, 4, true>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 16: return launch_confusion<EP, FP, 8, false>(pred, truth, n, n_seg, seg, counts, pa, st);
        case 17: return launch_confusion<EP, FP, 8, true>(pred, truth, n, n_seg, seg, counts, pa, st);
    }
    set_error("unsupported truth element size %d (float=%d)", et, ft);
    return RFI_E_INVALID;
}

static int dispatch(const void* pred, int ep, int fpf, const void* truth, int et, int ft, long long n,
                    long long n_seg, long long seg, unsigned long long* counts, cudaStream_t st,
                    const PeerArgs& pa = PeerArgs{{nullptr}, 0, 0, 0ull}) {
    if (!counts) { set_error("counts is NULL"); return RFI_E_INVALID; }
    if ((n == 0 && pa.world == 0) || n_seg == 0) return RFI_OK;  // a rank with an empty shard still takes part in the exchange
    if (n != 0 && (!pred || !truth)) { set_error("pred / truth is NULL"); return RFI_E_INVALID; }
    switch (ep * 2 + (fpf ? 1 : 0)) {
        case 2:  return dispatch_true<1, false>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 4:  return dispatch_true<2, false>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 5:  return dispatch_true<2, true>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 8:  return dispatch_true<4, false>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 9:  return dispatch_true<4, true>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 16: return dispatch_true<8, false>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
        case 17: return dispatch_true<8, true>(pred, truth, et, ft, n, n_seg, seg, counts, pa, st);
    }
    set_error("unsupported pred element size %d (float=%d)", ep, fpf);
    return RFI_E_INVALID;
}

}  // namespace rfi

extern "C" int rfi_confusion_counts(const void* pred, int elem_pred, int is_float_pred, const void* truth,
                                    int elem_true, int is_float_true, int64_t n, unsigned long long* counts,
                                    void* stream) {
    if (n < 0) { rfi::set_error("n < 0"); return RFI_E_INVALID; }
    return rfi::dispatch(pred, elem_pred, is_float_pred, truth, elem_true, is_float_true, n, -1, 0, counts,
                         (cudaStream_t)stream);
}

extern "C" int rfi_confusion_counts_allreduce(const void* pred, int elem_pred, int is_float_pred,
                                              const void* truth, int elem_true, int is_float_true, int64_t n,
                                              void* const* peers, int world, int rank, uint64_t epoch,
                                              unsigned long long* counts, void* stream) {
    if (n < 0) { rfi::set_error("n < 0"); return RFI_E_INVALID; }
    if (!peers || world < 1 || world > rfi::kPeerMax || rank < 0 || rank >= world || epoch == 0) {
        rfi::set_error("bad peer arguments (world 1..%d, 0 <= rank < world, epoch >= 1)", rfi::kPeerMax);
        return RFI_E_INVALID;
    }
    rfi::PeerArgs pa;
    for (int i = 0; i < rfi::kPeerMax; ++i) pa.peer[i] = i < world ? static_cast<rfi::PeerBuf*>(peers[i]) : nullptr;
    for (int i = 0; i < world; ++i) if (!pa.peer[i]) { rfi::set_error("peer %d is NULL", i); return RFI_E_INVALID; }
    pa.world = world; pa.rank = rank; pa.epoch = epoch;
    return rfi::dispatch(pred, elem_pred, is_float_pred, truth, elem_true, is_float_true, n, -1, 0, counts,
                         (cudaStream_t)stream, pa);
}

// exchange buffer of one rank: its own allocation (not the caller's caching allocator) so that the
// CUDA IPC handle maps exactly this buffer in the peers
extern "C" int rfi_peer_alloc(void** buf, unsigned char* handle64) {
    if (!buf || !handle64) { rfi::set_error("NULL argument"); return RFI_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    static_assert(sizeof(rfi::PeerBuf) <= RFI_PEER_BYTES, "exchange buffer size");
    RFI_CUDA_TRY(cudaMalloc(buf, 2u << 20));
    RFI_CUDA_TRY(cudaMemset(*buf, 0, 2u << 20));
    RFI_CUDA_TRY(cudaDeviceSynchronize());
    RFI_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *buf));
    return RFI_OK;
}
extern "C" int rfi_peer_open(const unsigned char* handle64, void** peer) {
    if (!peer || !handle64) { rfi::set_error("NULL argument"); return RFI_E_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    RFI_CUDA_TRY(cudaIpcOpenMemHandle(peer, h, cudaIpcMemLazyEnablePeerAccess));
    return RFI_OK;
}
extern "C" int rfi_peer_close(void* peer) {
    if (peer) RFI_CUDA_TRY(cudaIpcCloseMemHandle(peer));
    return RFI_OK;
}
extern "C" int rfi_peer_free(void* buf) {
    if (buf) RFI_CUDA_TRY(cudaFree(buf));
    return RFI_OK;
}

extern "C" int rfi_confusion_counts_segmented(const void* pred, int elem_pred, int is_float_pred,
                                              const void* truth, int elem_true, int is_float_true,
                                              int64_t n_seg, int64_t seg, unsigned long long* counts,
                                              void* stream) {
    if (n_seg < 0 || seg < 0) { rfi::set_error("negative segment count / size"); return RFI_E_INVALID; }
    if (seg == 0) {
        if (n_seg && counts) {
            cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 3 * n_seg, (cudaStream_t)stream);
            if (e != cudaSuccess) return rfi::cuda_fail(e, "cudaMemsetAsync");
        }
        return RFI_OK;
    }
    return rfi::dispatch(pred, elem_pred, is_float_pred, truth, elem_true, is_float_true, n_seg * seg, n_seg,
                         seg, counts, (cudaStream_t)stream);
}

// ---- self-test hook (tests only): sqrt_rn_unit vs sqrt.rn.f32 over [1, 2] ---------------------
#include "rfi_tiles.cuh"
namespace rfi {
__global__ void sqrt_unit_check_kernel(unsigned long long* mismatches) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 2^23
    if (i > (1u << 23)) return;
    const float t = __uint_as_float(0x3f800000u + i);
    if (__float_as_uint(sqrt_rn_unit(t)) != __float_as_uint(__fsqrt_rn(t))) atomicAdd(mismatches, 1ull);
}
// cabs_fast (own quotient sequence, one range test) vs cabs_np (div.rn, sqrt.rn) on 2^32 operand pairs:
// a third with arbitrary bit patterns (zeros, denormals, inf, NaN, any exponent gap), a third with both
// operands inside a few binades of each other at an arbitrary exponent, a third around the edges of the
// fast range; and the quotient itself against div.rn wherever cabs depends on it (quotient >= 2^-13).
__global__ void cabs_fast_check_kernel(unsigned long long* mismatches) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    for (uint32_t it = 0; it < 4096u; ++it) {
        uint32_t x = t * 4096u + it, y;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;   // two hashed words
        y = x + 0x9e3779b9u; y ^= y >> 16; y *= 0x85ebca6bu; y ^= y >> 13; y *= 0xc2b2ae35u; y ^= y >> 16;
        uint32_t a = x, b = y;
        const uint32_t kind = it % 3u;
        if (kind == 1u) {          // close exponents
            const uint32_t e = 1u + (x >> 24) % 253u, gap = (y >> 28) & 15u;
            a = (x & 0x807fffffu) | (e << 23);
            b = (y & 0x807fffffu) | ((e > gap ? e - gap : 1u) << 23);
        } else if (kind == 2u) {   // edges of the fast range: 2^-64 and 2^64 (+-2 binades)
            const uint32_t e = ((x >> 24) & 1u ? 63u : 191u) - 2u + ((x >> 25) & 3u), gap = (y >> 27) & 31u;
            a = (x & 0x807fffffu) | (e << 23);
            b = (y & 0x807fffffu) | ((e > gap ? e - gap : 0u) << 23);
        }
        if (x & 0x00800000u) { const uint32_t s = a; a = b; b = s; }
        const float re = __uint_as_float(a), im = __uint_as_float(b);
        const float f = cabs_fast(re, im), g = cabs_np<float>(re, im);
        const bool same = (f != f && g != g) || __float_as_uint(f) == __float_as_uint(g);
        bad += same ? 0u : 1u;
        const uint32_t u = a & 0x7fffffffu, v = b & 0x7fffffffu;
        const uint32_t hb = u > v ? u : v, lb = u > v ? v : u;
        if (hb - kCabsLoBits < kCabsSpanBits) {
            const float hi = __uint_as_float(hb), lo = __uint_as_float(lb);
            const float q = __fdiv_rn(lo, hi);
            if (q >= 0.0001220703125f && __float_as_uint(div_rn_unit(lo, hi)) != __float_as_uint(q)) ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}
}  // namespace rfi
extern "C" int rfi_selftest_cabs_fast(unsigned long long* mismatches_dev, void* stream) {
    using namespace rfi;
    cabs_fast_check_kernel<<<(1u << 20) / 256, 256, 0, (cudaStream_t)stream>>>(mismatches_dev);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
extern "C" int rfi_selftest_sqrt_unit(unsigned long long* mismatches_dev, void* stream) {
    using namespace rfi;
    sqrt_unit_check_kernel<<<((1u << 23) + 256) / 256 + 1, 256, 0, (cudaStream_t)stream>>>(mismatches_dev);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
