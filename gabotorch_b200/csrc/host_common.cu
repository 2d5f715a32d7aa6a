// Host-side plumbing shared by every C-ABI entry: error string, launch check, device properties.
#include <stdarg.h>

#include "common.cuh"

namespace gabo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, static_cast<int>(e), cudaGetErrorString(e));
        return GABO_E_CUDA;
    }
    return GABO_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace gabo

extern "C" int gabo_version(void) { return GABO_VERSION; }
extern "C" const char* gabo_last_error(void) { return gabo::g_err; }
