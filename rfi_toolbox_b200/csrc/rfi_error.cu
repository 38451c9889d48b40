// rfi_error.cu -- per-thread error text behind rfi_last_error_string().
#include <cstdarg>
#include <cstdio>

#include "rfi_common.cuh"

namespace rfi {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RFI_E_CUDA;
}
}  // namespace rfi

extern "C" const char* rfi_last_error_string(void) { return rfi::g_err; }
extern "C" int rfi_abi_version(void) { return RFI_B200_ABI_VERSION; }
