// nbg_common.cuh -- shared device/host helpers for the sm_100a kernels.
//
// Everything here is HBM-bound streaming machinery: 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, SASS: UBLKCP) between global memory and shared-memory row
// tiles, conflict-free "odd chunk" shared-memory indexing, and error plumbing for the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/nbg_b200.h"

namespace nbg {

// ------------------------------------------------------------------------------ host side
extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return NBG_ERR_CUDA;
    }
    return NBG_OK;
}

inline int check_cuda(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return NBG_ERR_CUDA;
    }
    return NBG_OK;
}

constexpr int kNumSMs = 148;  // B200
constexpr size_t kMaxSmem = 200 * 1024;  // dynamic shared memory the tile kernels size themselves for
constexpr size_t kMaxSmemOptIn = 226 * 1024;  // what kernels are opted into (227 KB per CTA minus room for static smem)

// Opt a kernel into kMaxSmem bytes of dynamic shared memory, once per (kernel, device).
int allow_big_smem_impl(const void *kern, const char *what);
template <typename K>
inline int allow_big_smem(K kern, const char *what) {
    return allow_big_smem_impl(reinterpret_cast<const void *>(kern), what);
}

// ---------------------------------------------------------------------------- device side
template <typename T>
__device__ __forceinline__ bool is_nan(T x) {
    return x != x;
}
template <>
__device__ __forceinline__ bool is_nan<int32_t>(int32_t) {
    return false;
}
template <>
__device__ __forceinline__ bool is_nan<int64_t>(int64_t) {
    return false;
}

template <typename T>
__device__ __forceinline__ T quiet_nan();
template <>
__device__ __forceinline__ float quiet_nan<float>() {
    return __int_as_float(0x7fc00000);
}
template <>
__device__ __forceinline__ double quiet_nan<double>() {
    return __longlong_as_double(0x7ff8000000000000LL);
}

// Non-contracted double arithmetic: the reference is compiled without FMA contraction
// (numbagg/decorators.py:34-49), so `s *= d; s += x` must round twice here as well.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// ---- double division through the hardware reciprocal seed -----------------------------------
// An IEEE double division is ~20 instructions on the FP64 pipe, and the read-outs below divide up
// to four times per output -- more than the whole recurrence.  Here 1/b comes from the
// MUFU.RCP64H seed (2^-20) and two Newton steps, and every quotient gets the residual correction
// q' = q + (a - b*q) * y, which yields the correctly rounded a/b (Markstein): the read-outs contain
// SIGN GATES on exact cancellations (bias > 0 after one observation, var1*var2 > 0 for repeated
// values), so quotients must round exactly as the reference's divisions do.  3 instructions per
// quotient + 5 per distinct divisor.  Divisors outside [1e-280, 1e280] (zero, the subnormal tail of
// a long NaN run, inf, NaN) take the IEEE path so that x/0, 0/0 and inf behave as in the reference.
__device__ __forceinline__ double fast_rcp(double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    y = fma(y, fma(-b, y, 1.0), y);
    y = fma(y, fma(-b, y, 1.0), y);
    return y;
}
__device__ __forceinline__ bool rcp_ok(double b) {
    const double ab = fabs(b);
    return ab > 1e-280 && ab < 1e280;
}
// The IEEE division behind the rare cases, OUT OF LINE: written as `cond ? fast : a / b` the compiler
// if-converts the select and evaluates the division's own inline fast path (MUFU.RCP64H + 7 FP64
// instructions) next to ours for every element (seen in the SASS of the exp scans: two Newton chains
// per quotient).
static __device__ __noinline__ double ieee_div(double a, double b) { return a / b; }
static __device__ __noinline__ double ieee_sqrt(double v) { return sqrt(v); }
static __device__ __noinline__ double ieee_rsqrt(double v) { return rsqrt(v); }


// sqrt through the float reciprocal-sqrt seed (MUFU.RSQ, ~2^-22 relative error), two Newton
// steps on y ~ 1/sqrt(v) and one correction of r = v*y: <= 1 ulp.  Values outside
// (1e-35, 1e35) -- including 0, negatives, inf and NaN -- take the IEEE sqrt path.
__device__ __forceinline__ double rsqrt_seed(double v) {
    return (double)rsqrtf((float)v);  // MUFU.RSQ on the float image of v (callers bound v)
}
__device__ __forceinline__ double fast_rsqrt(double v) {
    // callers guarantee 2^-120 < v < 2^120 or handle the specials themselves
    double y = rsqrt_seed(v);                 // ~2^-22
    double h = 0.5 * v;
    y = y * fma(-h * y, y, 1.5);              // ~2^-43
    y = y * fma(-h * y, y, 1.5);              // ~2^-86 -> rounding-limited
    return y;
}
__device__ __forceinline__ double fast_sqrt(double v) {
    const bool tiny_or_huge = !(v > 1e-35 && v < 1e35);
    if (tiny_or_huge) return ieee_sqrt(v);  // rare (also NaN / negative / 0 / inf): IEEE path, out of line
    const double y = fast_rsqrt(v);
    double r = v * y;
    r = fma(fma(-r, r, v), 0.5 * y, r);
    return r;
}
// a / b given y = fast_rcp(b)
__device__ __forceinline__ double qdiv(double a, double b, double y) {
    const double q = a * y;
    const double qc = fma(fma(-b, q, a), y, q);
    // an infinite / NaN numerator (or an overflowing quotient) turns the residual into NaN: IEEE path
    if (fabs(qc) < __longlong_as_double(0x7ff0000000000000LL)) return qc;
    return ieee_div(a, b);
}
__device__ __forceinline__ double fdiv(double a, double b) {
    // straight-line fast path, ONE test, one IEEE fallback
    const double y = fast_rcp(b);
    const double q = a * y;
    const double qc = fma(fma(-b, q, a), y, q);
    if (rcp_ok(b) && fabs(qc) < __longlong_as_double(0x7ff0000000000000LL)) return qc;
    return ieee_div(a, b);
}


// Product of two inputs in the INPUT type, then widened: numba types float32*float32 as
// float32 (verified bit-for-bit against the reference in tests/test_gpu_golden.py).
__device__ __forceinline__ double prod_as_input(float a, float b) { return (double)__fmul_rn(a, b); }
__device__ __forceinline__ double prod_as_input(double a, double b) { return __dmul_rn(a, b); }

// ----------------------------------------------------------------------------- key encoding
// Order-preserving map double -> u64; never 0 for a non-NaN input, so 0 can mean "empty".
__device__ __forceinline__ unsigned long long order_key(double v) {
    if (v == 0.0) v = 0.0;  // -0.0 and +0.0 compare equal in the reference
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copies (TMA engine, no tensor map needed) ----------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// plain arrival (release.cta): a consumer warp hands a ring stage back to the producer
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of a global range (no shared-memory destination): keeps DRAM reads in flight for
// tiles that the NEXT wave of CTAs will stage, independent of CTA occupancy
__device__ __forceinline__ void bulk_prefetch_l2(const void *gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, completion signalled on an mbarrier (bytes: multiple of 16, both
// addresses 16-byte aligned).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a store
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- row-span staging ------------------------------------------------------------------
// A "span" is the logical position range [p0, p0 + len) of one row of length n.  Positions
// outside [0, n) hold `fill`.  The shared buffer is addressed as s[j], j = p - p0, and is
// placed so that s has the same 16-byte phase as the global address of position p0:
// then every 16-byte aligned global address maps to a 16-byte aligned shared address and
// the whole 16-byte aligned middle of the span moves with ONE bulk copy; only the (< 16 B)
// unaligned head/tail and out-of-range positions are touched by threads.
template <typename T>
struct SpanPlan {
    int lo, hi;        // in-range part of the span as indices j in [0, len)
    int blk_lo;        // first j moved by the bulk copy
    uint32_t blk_bytes;  // bytes moved by the bulk copy (multiple of 16, may be 0)
};

template <typename T>
__device__ __forceinline__ uint32_t span_phase(const T *row, int64_t p0) {
    // address of position p0 (may lie before the row start) modulo 16
    int64_t addr = (int64_t)(uintptr_t)row + p0 * (int64_t)sizeof(T);
    return (uint32_t)(addr & 15);
}

template <typename T>
__device__ __forceinline__ SpanPlan<T> span_plan(const T *row, int64_t p0, int len, int64_t n) {
    SpanPlan<T> pl;
    int64_t lo = p0 < 0 ? -p0 : 0;
    int64_t hi = (p0 + len > n) ? (n - p0) : len;
    if (lo > len) lo = len;
    if (hi < lo) hi = lo;
    pl.lo = (int)lo;
    pl.hi = (int)hi;
    constexpr int PER16 = 16 / (int)sizeof(T);
    uintptr_t a_lo = (uintptr_t)(row + p0 + lo);
    int head = (int)(((16 - (a_lo & 15)) & 15) / sizeof(T));
    int cnt = pl.hi - pl.lo;
    if (head > cnt) head = cnt;
    int blk = ((cnt - head) / PER16) * PER16;
    pl.blk_lo = pl.lo + head;
    pl.blk_bytes = (uint32_t)blk * (uint32_t)sizeof(T);
    return pl;
}

// Threads fill everything the bulk copy does not: out-of-range positions and the unaligned
// head / tail.  (All threads of the CTA call this; `halo`, when non-null, supplies
// positions p < 0: halo[halo_len + p], for core-axis shards.)
template <typename T, int THREADS>
__device__ __forceinline__ void span_fill_edges(T *s, const T *row, int64_t p0, int len, const SpanPlan<T> &pl,
                                                T fill, const T *halo_row, int64_t halo_len) {
    const int tid = threadIdx.x;
    for (int j = tid; j < pl.lo; j += THREADS) {
        T v = fill;
        if (halo_row != nullptr) {
            int64_t h = halo_len + (p0 + j);
            if (h >= 0) v = halo_row[h];
        }
        s[j] = v;
    }
    for (int j = pl.lo + tid; j < pl.blk_lo; j += THREADS) s[j] = row[p0 + j];
    const int blk_hi = pl.blk_lo + (int)(pl.blk_bytes / sizeof(T));
    for (int j = blk_hi + tid; j < pl.hi; j += THREADS) s[j] = row[p0 + j];
    for (int j = pl.hi + tid; j < len; j += THREADS) s[j] = fill;
}

// L2 prefetch of the in-range, 16-byte aligned interior of a span (what a later CTA will stage).
template <typename T>
__device__ __forceinline__ void span_prefetch_l2(const T *row, int64_t p0, int len, int64_t n) {
    const int64_t lo = p0 < 0 ? 0 : p0;
    const int64_t hi = (p0 + len > n) ? n : (p0 + len);
    if (hi <= lo) return;
    const uintptr_t a = ((uintptr_t)(row + lo) + 15) & ~(uintptr_t)15;
    const uintptr_t b = (uintptr_t)(row + hi) & ~(uintptr_t)15;
    if (b > a) bulk_prefetch_l2(reinterpret_cast<const void *>(a), (uint32_t)(b - a));
}

// Tiles ahead of the running one that a CTA prefetches into L2 (0 = off).  Default: ~6.5 MB of
// input ahead, at most one wave of resident CTAs; NBG_PREFETCH_TILES overrides (tuning / A-B measurement).
int prefetch_distance(int resident_ctas_per_sm, size_t tile_bytes);

// Bulk-store s[0, cnt) to g[0, cnt): when s and g share their 16-byte phase thread 0 issues
// the aligned middle as one bulk store and all threads store the unaligned edges; otherwise
// (an output tensor whose rows are phased differently from the input's) every thread stores
// scalars.  Must be followed by bulk_wait_read_all() on thread 0 before the CTA reuses or
// releases the buffer.
template <typename T, int THREADS>
__device__ __forceinline__ void span_store(T *g, const T *s, int cnt) {
    constexpr int PER16 = 16 / (int)sizeof(T);
    const int tid = threadIdx.x;
    if ((((uintptr_t)g) & 15) != (smem_u32(s) & 15)) {
        for (int j = tid; j < cnt; j += THREADS) g[j] = s[j];
        return;
    }
    int head = (int)(((16 - ((uintptr_t)g & 15)) & 15) / sizeof(T));
    if (head > cnt) head = cnt;
    int blk = ((cnt - head) / PER16) * PER16;
    if (tid == 0 && blk > 0) {
        bulk_s2g(g + head, s + head, (uint32_t)blk * (uint32_t)sizeof(T));
        bulk_commit();
    }
    for (int j = tid; j < head; j += THREADS) g[j] = s[j];
    for (int j = head + blk + tid; j < cnt; j += THREADS) g[j] = s[j];
}

}  // namespace nbg
