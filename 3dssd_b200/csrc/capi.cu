// capi.cu -- status / error plumbing of the C ABI (include/ssd3d.h).
#include <stdarg.h>

#include "common.cuh"

namespace ssd3d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

}  // namespace ssd3d

extern "C" int ssd3d_version(void) { return 2; }
extern "C" const char *ssd3d_last_error(void) { return ssd3d::g_err; }
