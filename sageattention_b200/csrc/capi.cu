// C-ABI plumbing shared by all entry points: thread-local error text, device check, version.
#include "common.cuh"

namespace sab {
char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace sab

extern "C" const char* sab_last_error(void) { return sab::error_buffer(); }

extern "C" int sab_version(void) { return 100; }

extern "C" int sab_check_device(void) {
  int dev = 0;
  SAB_CUDA_OK(cudaGetDevice(&dev));
  static int cached_dev = -1, cached_major = 0;
  if (cached_dev != dev) {
    int major = 0;
    SAB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    cached_major = major;
    cached_dev = dev;
  }
  SAB_REQUIRE(cached_major == 10, SAB_ERR_ARCH,
              "sageattention_b200 needs a compute-capability 10.x (sm_100a, B200) device; found major %d. "
              "There is no fallback path.", cached_major);
  return SAB_OK;
}
