// rfi_raw.cu -- raw complex patches for GPUPreprocessor.create_raw_patches
// (rfi_toolbox/preprocessing/preprocessor.py:846-940, :942-972):
// non-overlapping P x P tiles of every waterfall (remainders dropped, patchify without padding,
// :22-42) or the whole waterfall when it is no larger than the patch (:885-890), the matching
// mask tiles (the caller's flags, or |z| > 0 when none are given, :881-883), blank tiles dropped
// (:904-913) and the survivors written ONCE at their final shuffled position -- a tiled gather,
// 2 x (element + 1) bytes of HBM traffic per kept sample.
#include "rfi_common.cuh"

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
