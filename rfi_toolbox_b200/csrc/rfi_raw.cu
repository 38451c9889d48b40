// rfi_raw.cu -- raw complex patches for GPUPreprocessor.create_raw_patches
// (rfi_toolbox/preprocessing/preprocessor.py:846-940, :942-972):
// non-overlapping P x P tiles of every waterfall (remainders dropped, patchify without padding,
// :22-42) or the whole waterfall when it is no larger than the patch (:885-890), the matching
// mask tiles (the caller's flags, or |z| > 0 when none are given, :881-883), blank tiles dropped
// (:904-913) and the survivors written ONCE at their final shuffled position -- a tiled gather,
// 2 x (element + 1) bytes of HBM traffic per kept sample.
#include <string.h>

#include "rfi_tiles.cuh"

namespace rfi {

struct RawGeom {
    long long C, T, n_waterfalls;
    int Pr, Pc;       // tile rows / cols (P, P; or C, T when patchify is skipped)
    int nh, nw, per;  // tiles per waterfall
    int esize;        // bytes per sample: 8 (complex64) or 16 (complex128)
};

// |z| > 0 as NumPy evaluates it: False for 0 and for NaN magnitudes, True for inf
template <typename T>
RFI_DEVINL bool mag_positive(T re, T im) { return cabs_np<T>(re, im) > T(0); }

RFI_DEVINL bool sample_flag(const RawGeom& g, const void* data, const uint8_t* flags, size_t idx) {
    if (flags) return flags[idx] != 0;
    if (g.esize == 8) { const float2 z = static_cast<const float2*>(data)[idx]; return mag_positive<float>(z.x, z.y); }
    const double2 z = static_cast<const double2*>(data)[idx];
    return mag_positive<double>(z.x, z.y);
}

constexpr int kRawThreads = 256;

// one CTA per tile: number of flagged samples (mask.any() of preprocessor.py:906)
__global__ void __launch_bounds__(kRawThreads)
raw_count_kernel(RawGeom g, const void* __restrict__ data, const uint8_t* __restrict__ flags, int* __restrict__ counts) {
    const long long tile = blockIdx.x;
    const long long w = tile / g.per;
    const int t = (int)(tile % g.per), ti = t / g.nw, tj = t % g.nw;
    const size_t origin = ((size_t)w * g.C + (size_t)ti * g.Pr) * g.T + (size_t)tj * g.Pc;
    unsigned n = 0;
    const long long total = (long long)g.Pr * g.Pc;
    for (long long e = threadIdx.x; e < total; e += kRawThreads) {
        const long long r = e / g.Pc, c = e - r * g.Pc;
        n += sample_flag(g, data, flags, origin + (size_t)r * g.T + c) ? 1u : 0u;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    __shared__ unsigned tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&tot, n);
    __syncthreads();
    if (threadIdx.x == 0) counts[tile] = (int)tot;
}

// grid (tiles, row chunks): tile -> its slot; one sample (8 / 16 bytes) per thread per step, rows contiguous
template <int ESIZE>
__global__ void __launch_bounds__(kRawThreads)
raw_gather_kernel(RawGeom g, const void* __restrict__ data, const uint8_t* __restrict__ flags,
                  const long long* __restrict__ dest_slot, void* __restrict__ patches, uint8_t* __restrict__ masks) {
    const long long tile = blockIdx.x;
    const long long slot = dest_slot[tile];
    if (slot < 0) return;
    const long long w = tile / g.per;
    const int t = (int)(tile % g.per), ti = t / g.nw, tj = t % g.nw;
    const size_t origin = ((size_t)w * g.C + (size_t)ti * g.Pr) * g.T + (size_t)tj * g.Pc;
    const size_t out0 = (size_t)slot * g.Pr * g.Pc;
    using V = typename std::conditional<ESIZE == 8, float2, double2>::type;
    const V* src = static_cast<const V*>(data);
    V* dst = static_cast<V*>(patches);
    const int rows_per_cta = (g.Pr + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(g.Pr, r0 + rows_per_cta);
    for (int r = r0; r < r1; ++r) {
        for (int c = threadIdx.x; c < g.Pc; c += kRawThreads) {
            const size_t in = origin + (size_t)r * g.T + c;
            const V z = src[in];
            dst[out0 + (size_t)r * g.Pc + c] = z;
            bool f;
            if (flags) f = flags[in] != 0;
            else if constexpr (ESIZE == 8) f = mag_positive<float>(z.x, z.y);
            else f = mag_positive<double>(z.x, z.y);
            masks[out0 + (size_t)r * g.Pc + c] = f ? 1 : 0;
        }
    }
}

static int make_raw(int dtype, int64_t n_waterfalls, int64_t C, int64_t T, int P, RawGeom& g) {
    if (dtype != RFI_C64 && dtype != RFI_C128) { set_error("raw patches need complex input (preprocessor.py:832-836)"); return RFI_E_INVALID; }
    if (n_waterfalls < 0 || C <= 0 || T <= 0 || P <= 0) { set_error("bad shape / patch size"); return RFI_E_INVALID; }
    g.C = C; g.T = T; g.n_waterfalls = n_waterfalls; g.esize = dtype == RFI_C64 ? 8 : 16;
    if (C <= P && T <= P) { g.Pr = (int)C; g.Pc = (int)T; g.nh = g.nw = 1; }   // :885-890
    else { g.Pr = g.Pc = P; g.nh = (int)(C / P); g.nw = (int)(T / P); }        // patchify, step = P, no padding
    g.per = g.nh * g.nw;
    if ((long long)g.per * n_waterfalls > 0x7fffffffLL) { set_error("too many tiles"); return RFI_E_UNSUPPORTED; }
    return RFI_OK;
}

}  // namespace rfi

using namespace rfi;

extern "C" int64_t rfi_raw_num_tiles(int dtype, int64_t n_waterfalls, int64_t channels, int64_t times, int32_t patch,
                                     int32_t* tile_rows, int32_t* tile_cols) {
    RawGeom g;
    if (make_raw(dtype, n_waterfalls, channels, times, patch, g)) return -1;
    if (tile_rows) *tile_rows = g.Pr;
    if (tile_cols) *tile_cols = g.Pc;
    return (int64_t)g.per * n_waterfalls;
}

extern "C" int rfi_raw_tile_counts(const void* data, int dtype, const uint8_t* flags, int64_t n_waterfalls,
                                   int64_t channels, int64_t times, int32_t patch, int32_t* counts, void* stream) {
    RawGeom g;
    int rc = make_raw(dtype, n_waterfalls, channels, times, patch, g);
    if (rc) return rc;
    const long long tiles = (long long)g.per * n_waterfalls;
    if (tiles == 0) return RFI_OK;
    if (!data || !counts) { set_error("data / counts is NULL"); return RFI_E_INVALID; }
    raw_count_kernel<<<(unsigned)tiles, kRawThreads, 0, (cudaStream_t)stream>>>(g, data, flags, counts);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

extern "C" int rfi_raw_gather(const void* data, int dtype, const uint8_t* flags, int64_t n_waterfalls,
                              int64_t channels, int64_t times, int32_t patch, const int64_t* dest_slot,
                              void* patches, uint8_t* masks, void* stream) {
    RawGeom g;
    int rc = make_raw(dtype, n_waterfalls, channels, times, patch, g);
    if (rc) return rc;
    const long long tiles = (long long)g.per * n_waterfalls;
    if (tiles == 0) return RFI_OK;
    if (!data || !dest_slot || !patches || !masks) { set_error("NULL buffer"); return RFI_E_INVALID; }
    const int chunks = g.Pr >= 64 ? (g.Pr + 31) / 32 : 1;   // 32 rows per CTA
    const long long* dest = reinterpret_cast<const long long*>(dest_slot);
    const dim3 grid((unsigned)tiles, (unsigned)chunks);
    if (g.esize == 8) raw_gather_kernel<8><<<grid, kRawThreads, 0, (cudaStream_t)stream>>>(g, data, flags, dest, patches, masks);
    else raw_gather_kernel<16><<<grid, kRawThreads, 0, (cudaStream_t)stream>>>(g, data, flags, dest, patches, masks);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

// ------------------------------------------------------------------------------------------------
// Rotated, zero-padded copies of the waterfalls (preprocessor.py:413-446 `_apply_rotations`, then the
// bottom / right zero pad of `_create_patches`, :527-550, which FOLLOWS the flip / transpose): view r
// of every waterfall, padded to multiples of P, so that a geometry whose dims are not multiples of P
// can run through the on-chip kernels as four single-view plans (every rotated patch is its own
// statistics group there).  r0: X, r1: X[::-1, :], r2: X.T, r3: X.T[::-1, :].
namespace rfi {

template <typename V>
__global__ void __launch_bounds__(256)
rotate_pad_kernel(const V* __restrict__ in, V* __restrict__ out, long long C, long long T, long long Cp, long long Tp, int r) {
    // out is (W, Cp, Tp); 32 x 32 tiles through shared memory so that the transposed views read AND
    // write rows
    __shared__ V tile[32][33];
    const long long w = blockIdx.z;
    const long long y0 = (long long)blockIdx.y * 32, x0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const V* src = in + (size_t)w * C * T;
    V* dst = out + (size_t)w * Cp * Tp;
    V zero;
    memset(&zero, 0, sizeof(V));
    if (r <= 1) {
        for (int k = ty; k < 32; k += 8) {
            const long long y = y0 + k, x = x0 + tx;
            if (y < Cp && x < Tp) {
                const long long sy = (r == 0) ? y : C - 1 - y;
                dst[(size_t)y * Tp + x] = (y < C && x < T) ? src[(size_t)sy * T + x] : zero;
            }
        }
        return;
    }
    // transposed views: out[y][x] = X[x][y'] with y' = y (r2) or T - 1 - y (r3); y < T, x < C
    for (int k = ty; k < 32; k += 8) {          // read a 32 x 32 block of X by rows of X
        const long long sx = x0 + k;            // row of X  (= output column)
        const long long oy = y0 + tx;           // output row
        const long long sy = (r == 2) ? oy : T - 1 - oy;  // column of X
        tile[k][tx] = (sx < C && oy < T) ? src[(size_t)sx * T + sy] : zero;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const long long y = y0 + k, x = x0 + tx;
        if (y < Cp && x < Tp) dst[(size_t)y * Tp + x] = tile[tx][k];
    }
}

}  // namespace rfi

extern "C" int rfi_rotate_pad(const void* in, void* out, int elem_bytes, int64_t n_waterfalls, int64_t channels,
                              int64_t times, int64_t out_rows, int64_t out_cols, int rotation, void* stream) {
    if (rotation < 0 || rotation > 3) { set_error("rotation must be 0..3"); return RFI_E_INVALID; }
    if (n_waterfalls < 0 || channels <= 0 || times <= 0) { set_error("bad shape"); return RFI_E_INVALID; }
    const int64_t need_r = rotation <= 1 ? channels : times, need_c = rotation <= 1 ? times : channels;
    if (out_rows < need_r || out_cols < need_c) { set_error("padded shape smaller than the rotated view"); return RFI_E_INVALID; }
    if (n_waterfalls == 0) return RFI_OK;
    if (!in || !out) { set_error("NULL buffer"); return RFI_E_INVALID; }
    if (n_waterfalls > 65535 || (out_rows + 31) / 32 > 65535) { set_error("too many waterfalls / rows per call"); return RFI_E_UNSUPPORTED; }
    const dim3 grid((unsigned)((out_cols + 31) / 32), (unsigned)((out_rows + 31) / 32), (unsigned)n_waterfalls);
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: rotate_pad_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), channels, times, out_rows, out_cols, rotation); break;
        case 4: rotate_pad_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), static_cast<float*>(out), channels, times, out_rows, out_cols, rotation); break;
        case 8: rotate_pad_kernel<float2><<<grid, 256, 0, st>>>(static_cast<const float2*>(in), static_cast<float2*>(out), channels, times, out_rows, out_cols, rotation); break;
        case 16: rotate_pad_kernel<double2><<<grid, 256, 0, st>>>(static_cast<const double2*>(in), static_cast<double2*>(out), channels, times, out_rows, out_cols, rotation); break;
        default: set_error("element size %d not supported", elem_bytes); return RFI_E_INVALID;
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}

// ------------------------------------------------------------------------------------------
// Preprocessor.patches (preprocessor.py:194, 272-311, 345-359): the PROCESSED patches in the dataset's
// final order -- normalised / stretched / inf-filled samples of the real branch, the raw complex samples
// of the complex branch.  The hot path never materialises them (phase 2 goes from the cube to the image
// channels); this kernel rebuilds them on demand from the cube, the tile statistics of phase 1 and the
// canonical index of every output patch, with the very operations of phase 1 (process_sample).
namespace rfi {

template <int DT, bool kComplexBranch>
__global__ void __launch_bounds__(256)
processed_patches_kernel(PlanDev p, int P, const void* __restrict__ data, const rfi_tile_stat_t* __restrict__ stats,
                         const long long* __restrict__ order, void* __restrict__ out) {
    using T = typename In<DT>::T;
    const long long k = blockIdx.x;
    const long long q = order[k];
    const int R = p.rotations, per = p.nh * p.nw;
    const long long w = q / ((long long)R * per);
    const int rem = (int)(q % ((long long)R * per)), r = rem / per, t = rem % per;
    int ti, tj;
    if (r == 0) { ti = t / p.nw; tj = t % p.nw; }
    else if (r == 1) { ti = p.nh - 1 - t / p.nw; tj = t % p.nw; }
    else if (r == 2) { tj = t / p.nh; ti = t % p.nh; }
    else { tj = p.nw - 1 - t / p.nh; ti = t % p.nh; }
    const size_t origin = ((size_t)w * p.channels + (size_t)ti * P) * p.times + (size_t)tj * P;
    const rfi_tile_stat_t st = stats[w * per + (long long)ti * p.nw + tj];
    const T mb = (T)st.median_before, fill = (T)st.inf_fill, ma = (T)st.median_after;
    const long long total = (long long)P * P;
    for (long long e = (long long)blockIdx.y * 256 + threadIdx.x; e < total; e += (long long)gridDim.y * 256) {
        const int orow = (int)(e / P), ocol = (int)(e % P);
        int sr, sc;
        if (r == 0) { sr = orow; sc = ocol; }
        else if (r == 1) { sr = P - 1 - orow; sc = ocol; }
        else if (r == 2) { sr = ocol; sc = orow; }
        else { sr = ocol; sc = P - 1 - orow; }
        const size_t idx = origin + (size_t)sr * p.times + sc;
        if constexpr (kComplexBranch) {
            using Z = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
            static_cast<Z*>(out)[k * total + e] = static_cast<const Z*>(data)[idx];
        } else {
            T a, ph;
            load1<DT, false>(data, idx, a, ph);
            static_cast<T*>(out)[k * total + e] = process_sample<T>(a, p, mb, fill, ma);
        }
    }
}

}  // namespace rfi

extern "C" int rfi_processed_patches(const rfi_plan_t* plan, const void* data, const rfi_tile_stat_t* stats,
                                     const int64_t* order, int64_t n_out, void* out, void* stream) {
    using namespace rfi;
    if (!plan || plan->patch <= 0) { set_error("plan is NULL"); return RFI_E_INVALID; }
    if (rfi_plan_path(plan) == RFI_PATH_GENERIC) {
        set_error("rfi_processed_patches: only for geometries whose dims are multiples of the patch size (on-chip paths)");
        return RFI_E_UNSUPPORTED;
    }
    if (n_out < 0 || (n_out > 0 && (!data || !stats || !order || !out))) { set_error("bad arguments to rfi_processed_patches"); return RFI_E_INVALID; }
    if (n_out == 0) return RFI_OK;
    const int P = plan->patch;
    PlanDev d;
    d.n_waterfalls = plan->n_waterfalls; d.channels = plan->channels; d.times = plan->times;
    d.nh = (int)(plan->channels / P); d.nw = (int)(plan->times / P);
    d.rotations = plan->rotations; d.stretch = plan->stretch; d.norm_before = plan->norm_before;
    d.norm_after = plan->norm_after; d.flag_mode = plan->flag_mode; d.magnitude = plan->magnitude; d.sigma = plan->sigma;
    const bool cb = plan->dtype >= RFI_C64 && !plan->magnitude;
    cudaStream_t st = (cudaStream_t)stream;
    const long long* ord = reinterpret_cast<const long long*>(order);
    const unsigned gy = (unsigned)((P * P + 256 * 16 - 1) / (256 * 16));
    for (long long k0 = 0; k0 < n_out; k0 += 0x7fffffffLL) {   // grid.x limit
        const long long nk = n_out - k0 < 0x7fffffffLL ? n_out - k0 : 0x7fffffffLL;
        const dim3 grid((unsigned)nk, gy);
        const size_t esz = (plan->dtype == RFI_F32 ? 4 : plan->dtype == RFI_F64 ? 8 : plan->dtype == RFI_C64 ? 8 : 16);
        const size_t osz = cb ? esz : (plan->dtype == RFI_F32 || plan->dtype == RFI_C64 ? 4 : 8);
        void* o = static_cast<char*>(out) + (size_t)k0 * P * P * osz;
        switch (plan->dtype) {
            case RFI_F32:  processed_patches_kernel<RFI_F32, false><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o); break;
            case RFI_F64:  processed_patches_kernel<RFI_F64, false><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o); break;
            case RFI_C64:
                if (cb) processed_patches_kernel<RFI_C64, true><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o);
                else processed_patches_kernel<RFI_C64, false><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o);
                break;
            default:
                if (cb) processed_patches_kernel<RFI_C128, true><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o);
                else processed_patches_kernel<RFI_C128, false><<<grid, 256, 0, st>>>(d, P, data, stats, ord + k0, o);
                break;
        }
    }
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
