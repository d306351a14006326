// nbg_abi.cu -- process-wide state of the C ABI (error string, launch counter, version).
#include "nbg_common.cuh"

#include <stdlib.h>

#include <mutex>
#include <unordered_set>

namespace nbg {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int allow_big_smem_impl(const void *kern, const char *what) {
    static std::mutex mu;
    static std::unordered_set<uint64_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t key = (uint64_t)(uintptr_t)kern ^ ((uint64_t)(dev + 1) << 56);
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(key)) return NBG_OK;
    // static + dynamic shared memory of a CTA may not exceed 227 KB
    cudaFuncAttributes fa;
    int rc = check_cuda(cudaFuncGetAttributes(&fa, kern), what);
    if (rc) return rc;
    size_t dyn = (size_t)227 * 1024 - fa.sharedSizeBytes;
    if (dyn > kMaxSmemOptIn) dyn = kMaxSmemOptIn;
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), what);
    if (rc) return rc;
    done.insert(key);
    return NBG_OK;
}
int prefetch_distance(int resident_ctas_per_sm) {
    if (const char *e = getenv("NBG_PREFETCH_TILES")) return atoi(e);
    return kNumSMs * resident_ctas_per_sm;
}
}  // namespace nbg

extern "C" int nbg_abi_version(void) { return NBG_ABI_VERSION; }
extern "C" const char *nbg_last_error(void) { return nbg::g_err; }
extern "C" int64_t nbg_launch_count(void) { return nbg::g_launches.load(); }
