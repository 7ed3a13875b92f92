// ABI bookkeeping: version, thread-local error string, device query.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace sdb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_error(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return SDB_ERR_CUDA;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}

}  // namespace sdb

extern "C" int sdb_abi_version(void) { return SDB_ABI_VERSION; }
extern "C" const char* sdb_last_error(void) { return sdb::g_err; }
