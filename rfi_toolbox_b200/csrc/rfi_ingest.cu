// rfi_ingest.cu -- complex128 / float64 cubes taken in as complex64 / float32 on the way to the path.
//
// The reference's loaders hand out complex128 (MSLoader.load, rfi_toolbox/io/ms_loader.py:202-238:
// `np.zeros([4, total_channels, num_times], dtype="complex128")`, stacked per baseline; the synthetic
// generator likewise, synthetic_generator.py:648), and the reference then computes in float64 until the
// `.astype(np.float32)` of preprocessor.py:376.  A caller who opts in (Preprocessor(...,
// compute_dtype="float32")) gets the float32 arithmetic of the benchmark path instead: every component
// is rounded once (round-to-nearest-even, what `ndarray.astype(np.complex64)` does), 24 B / px of traffic,
// after which the cube is half the size and takes the on-chip kernels at every patch size.
#include "rfi_common.cuh"

namespace rfi {

// n = number of float64 scalars (2 per complex sample); 128-bit loads, 64-bit stores
__global__ void __launch_bounds__(256)
downcast_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n2 = n >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(in) + i);
        reinterpret_cast<float2*>(out)[i] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = __double2float_rn(in[n - 1]);
}

}  // namespace rfi

extern "C" int rfi_downcast(const void* in, void* out, int dtype_in, int64_t n, void* stream) {
    using namespace rfi;
    if (dtype_in != RFI_F64 && dtype_in != RFI_C128) { set_error("rfi_downcast: input must be float64 / complex128 (dtype %d)", dtype_in); return RFI_E_INVALID; }
    if (n < 0 || (n > 0 && (!in || !out))) { set_error("bad arguments to rfi_downcast"); return RFI_E_INVALID; }
    if (n == 0) return RFI_OK;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) { set_error("rfi_downcast: unaligned buffers"); return RFI_E_INVALID; }
    const long long scalars = dtype_in == RFI_C128 ? 2 * (long long)n : (long long)n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (scalars / 2 + 256 * 4 - 1) / (256 * 4);
    const long long cap = (long long)sms * 8;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    downcast_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const double*>(in), static_cast<float*>(out), scalars);
    RFI_CUDA_TRY(cudaGetLastError());
    return RFI_OK;
}
