// nbg_group_partition.cuh -- per-element labels at high cardinality, multi-channel additive ops
// (group_nanmean / nanvar / nanstd, numbagg/grouped.py:7-27, 148-208): partition by label range, then
// shared-memory bins per range (north_star: "sort-by-label segmented reduce").
//
// Why: the atomic path (nbg_group.cu) resolves one global RED per element and channel in L2, and the L2
// atomic units retire ~144 G RED/s on a B200 whatever the table layout -- BASELINE config 5 (2e9
// float64 elements, 1e7 int64 labels) takes 12.5 ms per channel pass, 37 ms for nanvar.  Fewer global
// atomics per element need the elements of one label range together:
//   1. part_hist_kernel     labels only: 148 persistent CTAs, each counts the buckets (bucket = label >> 13:
//                           8192 labels) of its contiguous slice of the input -> count matrix [cta][bucket]
//   2. part_colscan / part_scan   per bucket the exclusive prefix over the CTAs, then the bucket bases
//   3. part_scatter_kernel  the same CTAs walk their slices again: every element with an in-range label
//                           goes to slot base[bucket] + (earlier slices) + (rank in this slice, a
//                           shared-memory cursor) as (double value, 13-bit local label).  No global atomics:
//                           a first version that reserved one run per chunk and bucket with a global atomic
//                           spent 75 ms in 6e8 returning atomics on 1221 addresses
//   4. part_reduce_kernel   one CTA per bucket: 8192 x (sum, [sum of squares], count) bins in shared
//                           memory, shared-memory atomics (the double ones are CAS loops, conflicts are
//                           rare), then a plain read-modify-write of the bucket's slice of the workspace
//                           planes (the bucket owns its labels)
// Traffic: 8 (labels) + 16 (values + labels) + 10 written + 10 read = 44 B per element against 3 x 16
// for the three channel passes, and no global atomics on the table.  The scratch (10 bytes per
// element + a few KB) comes from the device's stream-ordered pool; if that allocation fails the caller
// falls back to the channel passes.  Summation order inside a label is not deterministic (it is not on
// the atomic path either); sums are double.
#pragma once

#include "nbg_common.cuh"

namespace nbg {

// Open write streams of the scatter = CTAs x buckets x 2 arrays, one 32-byte sector each; they must all
// stay in L2 until their sectors are full.  592 CTAs x 4883 buckets (2048 labels each) = 2.9 M streams
// evicted partial sectors 5 x over (108 GB of DRAM writes for 20 GB of data, 159 ms); 148 x 1221 = 180 k
// streams (12 MB) combine in L2.
constexpr int kPartShift = 13;
constexpr int kPartBW = 1 << kPartShift;  // labels per bucket: 8192 x 20 bytes of bins, one reducing CTA per SM
constexpr int kPartMaxBuckets = 4096;     // shared-memory cursors of the scatter kernel (32 KB)
constexpr int kPartThreads = 1024;
constexpr int kPartPer = 8;               // elements per thread and chunk
constexpr int kPartChunk = kPartThreads * kPartPer;
constexpr int kPartCtas = 148;            // persistent CTAs of the histogram / scatter kernels: contiguous slices

// mat[cta * nb + b] = in-range labels of bucket b in the CTA's slice (NaN values are carried along and
// dropped by the reduce kernel: counting them here would need the values too)
template <typename L>
__global__ void __launch_bounds__(kPartThreads) part_hist_kernel(const L *__restrict__ labels, int64_t n, int64_t K, int nb,
                                                                 int64_t chunks_per_cta, unsigned *__restrict__ mat) {
    extern __shared__ unsigned part_h[];
    for (int b = threadIdx.x; b < nb; b += kPartThreads) part_h[b] = 0;
    __syncthreads();
    const int64_t c_lo = (int64_t)blockIdx.x * chunks_per_cta;
    for (int64_t c = c_lo; c < c_lo + chunks_per_cta; c++) {
        const int64_t c0 = c * kPartChunk;
        if (c0 >= n) break;
        long long lab[kPartPer];
#pragma unroll
        for (int k = 0; k < kPartPer; k++) {
            const int64_t i = c0 + (int64_t)k * kPartThreads + threadIdx.x;
            lab[k] = i < n ? (long long)__ldcs(labels + i) : -1ll;
        }
#pragma unroll
        for (int k = 0; k < kPartPer; k++)
            if (lab[k] >= 0 && lab[k] < K) atomicAdd(&part_h[(int)(lab[k] >> kPartShift)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += kPartThreads) mat[(size_t)blockIdx.x * nb + b] = part_h[b];
}

// per bucket: exclusive prefix over the CTAs (in place) and the bucket total
__global__ void part_colscan_kernel(unsigned *__restrict__ mat, unsigned long long *__restrict__ total, int nb, int nctas) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    unsigned long long run = 0;
    for (int c = 0; c < nctas; c++) {
        const unsigned v = mat[(size_t)c * nb + b];
        mat[(size_t)c * nb + b] = (unsigned)run;  // < 2^32: a bucket of one call holds fewer than 2^32 elements (n < 2^32)
        run += v;
    }
    total[b] = run;
}

// base[b] = first slot of bucket b, base[nb] = all slots (one CTA of 1024 threads, nb <= 4096 < 8192: 8 per thread)
__global__ void part_scan_kernel(const unsigned long long *__restrict__ total, unsigned long long *__restrict__ base, int nb) {
    __shared__ unsigned long long tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int b = tid * 8 + k;
        v[k] = b < nb ? total[b] : 0ull;
        s += v[k];
    }
    unsigned long long inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) tot[wid] = inc;
    __syncthreads();
    unsigned long long before = inc - s;
    for (int w = 0; w < wid; w++) before += tot[w];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int b = tid * 8 + k;
        if (b <= nb) base[b] = before;
        before += v[k];
    }
}

// every element with an in-range label goes to its bucket: slot = base[b] + (elements of b in earlier
// slices) + (rank inside this slice, from a shared-memory cursor) -- no global atomics
template <typename V, typename L>
__global__ void __launch_bounds__(kPartThreads) part_scatter_kernel(const V *__restrict__ values, const L *__restrict__ labels,
                                                                    int64_t n, int64_t K, int nb, int64_t chunks_per_cta,
                                                                    const unsigned *__restrict__ mat,
                                                                    const unsigned long long *__restrict__ base,
                                                                    double *__restrict__ pv, unsigned short *__restrict__ pl) {
    extern __shared__ __align__(8) unsigned char part_smem[];
    // first slot of this slice's run per bucket (64-bit) + a 32-bit rank counter: the 32-bit shared-memory
    // add is a native ATOMS.ADD, the 64-bit one a compare-and-swap spin loop (measured: 109 ms, issue slots
    // 3 % busy, all of it in the retry branch)
    unsigned long long *start = reinterpret_cast<unsigned long long *>(part_smem);  // [nb]
    unsigned *cur = reinterpret_cast<unsigned *>(start + nb);                          // [nb]
    for (int b = threadIdx.x; b < nb; b += kPartThreads) {
        start[b] = base[b] + mat[(size_t)blockIdx.x * nb + b];
        cur[b] = 0;
    }
    __syncthreads();
    const int64_t c_lo = (int64_t)blockIdx.x * chunks_per_cta;
    for (int64_t c = c_lo; c < c_lo + chunks_per_cta; c++) {
        const int64_t c0 = c * kPartChunk;
        if (c0 >= n) break;
        long long lab[kPartPer];
        V v[kPartPer];
#pragma unroll
        for (int k = 0; k < kPartPer; k++) {
            const int64_t i = c0 + (int64_t)k * kPartThreads + threadIdx.x;
            lab[k] = i < n ? (long long)__ldcs(labels + i) : -1ll;
            v[k] = i < n ? __ldcs(values + i) : (V)0;
        }
#pragma unroll
        for (int k = 0; k < kPartPer; k++) {
            if (lab[k] >= 0 && lab[k] < K) {
                const int b = (int)(lab[k] >> kPartShift);
                const unsigned long long pos = start[b] + atomicAdd(&cur[b], 1u);
                pv[pos] = (double)v[k];
                pl[pos] = (unsigned short)(lab[k] & (kPartBW - 1));
            }
        }
    }
}

// NSUM = 1: sum + count (mean); NSUM = 2: sum + sum of squares + count (var / std).  SQ_F32: the values
// are float32 images, their squares are rounded to float32 first (numba: V * V is V).
template <int NSUM, bool SQ_F32>
__global__ void __launch_bounds__(1024) part_reduce_kernel(const double *__restrict__ pv, const unsigned short *__restrict__ pl,
                                                          const unsigned long long *__restrict__ base, int64_t K,
                                                          double *__restrict__ out_sum, double *__restrict__ out_sq,
                                                          long long *__restrict__ out_cnt) {
    extern __shared__ __align__(8) unsigned char part_smem[];
    double *s0 = reinterpret_cast<double *>(part_smem);
    double *s1 = s0 + kPartBW;                                           // only when NSUM == 2
    int *cn = reinterpret_cast<int *>(s0 + (size_t)NSUM * kPartBW);
    const int tid = threadIdx.x, b = blockIdx.x;
    for (int l = tid; l < kPartBW; l += 1024) {
        s0[l] = 0.0;
        if (NSUM == 2) s1[l] = 0.0;
        cn[l] = 0;
    }
    __syncthreads();
    const unsigned long long lo = base[b], hi = base[b + 1];
    constexpr int U = 4;
    for (unsigned long long i0 = lo + tid; i0 < hi; i0 += (unsigned long long)U * 1024) {
        double x[U];
        int l[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long i = i0 + (unsigned long long)u * 1024;
            l[u] = i < hi ? (int)__ldcs(pl + i) : -1;
            x[u] = i < hi ? __ldcs(pv + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (l[u] >= 0 && x[u] == x[u]) {  // NaN values travelled with their labels: dropped here
                atomicAdd(&s0[l[u]], x[u]);
                if (NSUM == 2) atomicAdd(&s1[l[u]], SQ_F32 ? (double)__fmul_rn((float)x[u], (float)x[u]) : __dmul_rn(x[u], x[u]));
                atomicAdd(&cn[l[u]], 1);
            }
        }
    }
    __syncthreads();
    for (int l = tid; l < kPartBW; l += 1024) {
        const int64_t label = (int64_t)b * kPartBW + l;
        if (label < K && cn[l] != 0) {
            out_sum[label] += s0[l];
            if (NSUM == 2) out_sq[label] += s1[l];
            out_cnt[label] += (long long)cn[l];
        }
    }
}

// Returns NBG_OK with *handled = true when the partition path ran; *handled = false: not applicable
// (or no scratch memory) -- the caller runs the channel passes.
template <typename V, typename L, int NSUM>
static int launch_partition(const V *values, const L *labels, int64_t n, int64_t K, double *out_sum, double *out_sq,
                            long long *out_cnt, cudaStream_t stream, bool *handled) {
    *handled = false;
    const int64_t nb64 = (K + kPartBW - 1) >> kPartShift;
    if (n < ((int64_t)1 << 22) || n >= ((int64_t)1 << 32) || nb64 > kPartMaxBuckets || nb64 < 8) return NBG_OK;
    // EXPERIMENT, off unless NBG_GROUP_PARTITION=1: correct (tests/test_gpu_parity.py::test_group_partition_path)
    // but 3 x SLOWER than the channel passes on config 5 (123 vs 37 ms for nanvar): the histogram pass runs at
    // 7 TB/s (2.3 ms) and the reduce pass takes 10.5 ms, but the scatter pass takes 110 ms -- its per-element
    // RETURNING shared-memory atomic (the rank inside the bucket) is resolved one lane at a time (~490 cycles
    // per warp instruction with 32 distinct addresses; ncu: issue slots 3 % busy, short_scoreboard 181 cycles
    // per issue), where the non-returning adds of the histogram kernel cost nothing.  Ranking with warp-private
    // cursors and match.any instead (no atomics at all, 592 warp slices) was tried too: 135 ms -- so the pass is
    // really bound by its 4e9 single-element store transactions (32 distinct sectors per warp instruction:
    // ~36 G transactions/s); the missing piece is write-combining in shared memory (one 32-byte run per
    // bucket and CTA before anything is stored).
    const char *pe = getenv("NBG_GROUP_PARTITION");
    if (!pe || atoi(pe) == 0) return NBG_OK;
    const int nb = (int)nb64;
    const int64_t chunks = (n + kPartChunk - 1) / kPartChunk;
    const int nctas = (int)(chunks < kPartCtas ? chunks : kPartCtas);
    const int64_t cpc = (chunks + nctas - 1) / nctas;
    const size_t mat_bytes = ((size_t)nctas * nb * sizeof(unsigned) + 255) & ~(size_t)255;
    const size_t meta = (size_t)2 * (kPartMaxBuckets + 8) * sizeof(unsigned long long);
    const size_t pv_bytes = (size_t)n * sizeof(double), pl_bytes = ((size_t)n * sizeof(unsigned short) + 255) & ~(size_t)255;
    unsigned char *scratch = nullptr;
    {
        static bool pool_done[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !pool_done[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~(uint64_t)0;  // the scratch is 10 bytes per element: keep it cached between calls
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_done[dev] = true;
        }
    }
    if (cudaMallocAsync(reinterpret_cast<void **>(&scratch), meta + mat_bytes + pv_bytes + pl_bytes, stream) != cudaSuccess) {
        cudaGetLastError();  // not an error of this call: fall back
        return NBG_OK;
    }
    unsigned long long *total = reinterpret_cast<unsigned long long *>(scratch);
    unsigned long long *base = total + kPartMaxBuckets + 8;
    unsigned *mat = reinterpret_cast<unsigned *>(scratch + meta);
    double *pv = reinterpret_cast<double *>(scratch + meta + mat_bytes);
    unsigned short *pl = reinterpret_cast<unsigned short *>(scratch + meta + mat_bytes + pv_bytes);
    int rc = NBG_OK;
    part_hist_kernel<L><<<(unsigned)nctas, kPartThreads, (size_t)nb * sizeof(unsigned), stream>>>(labels, n, K, nb, cpc, mat);
    rc = check_launch("nbg_group(partition: histogram)");
    if (!rc) {
        part_colscan_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(mat, total, nb, nctas);
        rc = check_launch("nbg_group(partition: column scan)");
    }
    if (!rc) {
        part_scan_kernel<<<1, 1024, 0, stream>>>(total, base, nb);
        rc = check_launch("nbg_group(partition: scan)");
    }
    if (!rc) {
        auto kern = part_scatter_kernel<V, L>;
        const size_t smem = (size_t)nb * 12;
        if (smem > ((size_t)48 << 10)) rc = allow_big_smem(kern, "nbg_group(partition: scatter): cudaFuncSetAttribute");
        if (!rc) {
            kern<<<(unsigned)nctas, kPartThreads, smem, stream>>>(values, labels, n, K, nb, cpc, mat, base, pv, pl);
            rc = check_launch("nbg_group(partition: scatter)");
        }
    }
    if (!rc) {
        constexpr bool kF32 = std::is_same<V, float>::value;
        const size_t smem = (size_t)NSUM * kPartBW * sizeof(double) + (size_t)kPartBW * sizeof(int);
        auto kern = part_reduce_kernel<NSUM, kF32>;
        rc = allow_big_smem(kern, "nbg_group(partition: reduce): cudaFuncSetAttribute");
        if (!rc) {
            kern<<<(unsigned)nb, 1024, smem, stream>>>(pv, pl, base, K, out_sum, out_sq, out_cnt);
            rc = check_launch("nbg_group(partition: reduce)");
        }
    }
    cudaFreeAsync(scratch, stream);
    if (!rc) *handled = true;
    return rc;
}

}  // namespace nbg
