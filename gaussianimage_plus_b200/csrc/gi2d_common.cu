// gi2d_common.cu -- error reporting shared by all translation units of libgi2d.
#include <cstdarg>
#include <cstdio>
#include "gi2d_common.cuh"

namespace gi2d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// The reference never checks its launches (SURVEY 8b "Error convention"); we do, without
// synchronising: cudaPeekAtLastError reports launch-configuration failures immediately.
int check_launch(const char *what) {
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return GI2D_ERR_CUDA;
    }
    return GI2D_OK;
}

}  // namespace gi2d

extern "C" int gi2d_abi_version(void) { return GI2D_ABI_VERSION; }
extern "C" const char *gi2d_last_error(void) { return gi2d::g_err; }
