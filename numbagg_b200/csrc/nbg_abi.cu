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
int prefetch_distance(int resident_ctas_per_sm, size_t tile_bytes) {
    if (const char *e = getenv("NBG_PREFETCH_TILES")) return atoi(e);
    // ~6.5 MB ahead of the read front, at most one wave of resident CTAs.  Measured (round 2, config 3
    // ffill, 35 KB tiles): 150 tiles ahead 2.90 ms, 300: 2.99, 600: 3.06, one wave (888): 3.37 with 19.5 GB
    // of DRAM traffic for 16 GB of data -- lines prefetched too early are evicted by the output stream
    // before their tile is staged; off: 3.22 ms.
    const int wave = kNumSMs * resident_ctas_per_sm;
    int d = tile_bytes ? (int)(((size_t)13 << 19) / tile_bytes) : wave;
    if (d < 32) d = 32;
    return d < wave ? d : wave;
}
}  // namespace nbg

extern "C" int nbg_abi_version(void) { return NBG_ABI_VERSION; }
extern "C" const char *nbg_last_error(void) { return nbg::g_err; }
extern "C" int64_t nbg_launch_count(void) { return nbg::g_launches.load(); }
