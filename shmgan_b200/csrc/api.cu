// api.cu -- error plumbing and device queries of libshmgan.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void shm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int shm_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

extern "C" const char* shm_last_error(void) { return g_err; }
extern "C" int shm_version(void) { return 100; }
extern "C" int shm_sm_count(void) { return shm_num_sms(); }

// zero-fill of a caller-owned device buffer on the caller's stream (gradient buffers, loss seeds, accumulator arenas): one memset node instead of a
// framework fill kernel -- the step launches nothing but this library's kernels and memsets
extern "C" int shm_zero(void* ptr, int64_t bytes, void* stream) {
    SHM_REQUIRE(ptr != nullptr && bytes >= 0, "shm_zero: bad args");
    if (bytes == 0) return SHM_OK;
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) SHM_FAIL(SHM_ECUDA, "shm_zero: %s", cudaGetErrorString(e));
    return SHM_OK;
}
