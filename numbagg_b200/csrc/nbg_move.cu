// nbg_move.cu -- moving-window kernels (move_mean/sum/std/var/cov/corr) for sm_100a.
//
// Replaces the per-slice loops of numbagg/moving.py:12-275 (dispatched by ndmove,
// numbagg/decorators.py:275-341).
//
// Row-tile kernel (core axis contiguous, inner == 1)
//   One CTA owns T = THREADS*E consecutive outputs of one row.  It stages the input span
//   [c0 - window, c0 + T) in shared memory with ONE 1-D TMA bulk copy per input
//   (cp.async.bulk + mbarrier), each thread then owns E consecutive outputs and runs the
//   reference's running-window recurrence (subtract the element leaving, add the element
//   entering, NaN-skipping, integer valid-count) over them, reading the leading element at
//   j = window + t*E + k and the trailing one at j = t*E + k.  E is ODD, so both streams hit
//   32 distinct banks without padding and the tile stays a flat array the TMA engine can
//   fill and drain.  Results go to a shared out tile and leave with one bulk store.
//   The window state at a chunk start ("periodic re-sync") is rebuilt from the tile only:
//     window <= kDirectMax : sum the `window` preceding elements in order;
//     otherwise            : per-chunk totals -> block exclusive scan -> difference of two
//                            tile-local prefixes + (window mod E) elements in order.
//   Tile-local prefixes keep the rounding error at (T+window)/window ulps, far below the
//   reference's own drift over a whole row.  All sums are double, products
//   of float32 inputs are rounded to float32 first -- as numba types them.
//
// Column-walk kernel (inner > 1: any other axis of a C-contiguous array)
//   One thread per (outer, inner) column walks the core axis sequentially -- adjacent
//   threads touch adjacent addresses, so every access is coalesced; long columns are cut
//   into segments and each segment rebuilds its window from the `window` preceding elements.
//
// Algorithmic traffic: one read per input element + one write per output element.
#include <string.h>

#include "nbg_common.cuh"

namespace nbg {

constexpr int kDirectMax = 32;
constexpr int kRcpMax = 4096;  // windows up to this size use the shared reciprocal table

// (fast_rsqrt / fast_sqrt: nbg_common.cuh)
// ------------------------------------------------------------------------------------ ops
// Each op lists its channels (running double sums), how one observation contributes, the
// order of add/remove (moving.py differs between move_sum and the others) and the output.
template <typename T>
struct OpMean {
    static constexpr int NIN = 1, NCH = 1;
    static constexpr bool ADD_FIRST = false;  // moving.py:48-53 removes, then adds
    static constexpr int MIN_COUNT_FLOOR = 1;  // moving.py:18
    __device__ static __forceinline__ void contrib(T a, T, double *c) { c[0] = (double)a; }
    __device__ static __forceinline__ T finalize(const double *s, int count) {
        return (T)(s[0] / (double)count);
    }
    // rc = 1/count, rc1 = 1/(count-1) from the per-CTA reciprocal table (<= 1 ulp from a
    // true division; see DESIGN.md "finalisation")
    __device__ static __forceinline__ T finalize_fast(const double *s, double rc, double) { return (T)dmul(s[0], rc); }
    __device__ static __forceinline__ T finalize_pfx(const double *s, double rc, double rc1, bool &) { return finalize_fast(s, rc, rc1); }
};
template <typename T>
struct OpSum {
    static constexpr int NIN = 1, NCH = 1;
    static constexpr bool ADD_FIRST = true;  // moving.py:92-97 adds, then removes
    static constexpr int MIN_COUNT_FLOOR = 0;  // no clamp: min_count=0 -> 0.0 on empty windows
    __device__ static __forceinline__ void contrib(T a, T, double *c) { c[0] = (double)a; }
    __device__ static __forceinline__ T finalize(const double *s, int) { return (T)s[0]; }
    __device__ static __forceinline__ T finalize_fast(const double *s, double, double) { return (T)s[0]; }
    __device__ static __forceinline__ T finalize_pfx(const double *s, double rc, double rc1, bool &) { return finalize_fast(s, rc, rc1); }
};
template <typename T, bool SQRT>
struct OpVar {
    static constexpr int NIN = 1, NCH = 2;
    static constexpr bool ADD_FIRST = false;
    static constexpr int MIN_COUNT_FLOOR = 2;  // moving.py:127,158
    __device__ static __forceinline__ void contrib(T a, T, double *c) {
        c[0] = (double)a;
        c[1] = prod_as_input(a, a);
    }
    __device__ static __forceinline__ T finalize(const double *s, int count) {
        // (asum_sq - asum**2 / count) / (count - 1)   moving.py:146,178
        double v = dsub(s[1], dmul(s[0], s[0]) / (double)count) / (double)(count - 1);
        return (T)(SQRT ? sqrt(v) : v);
    }
    __device__ static __forceinline__ T finalize_fast(const double *s, double rc, double rc1) {
        const double v = dmul(dsub(s[1], dmul(dmul(s[0], s[0]), rc)), rc1);
        if constexpr (SQRT && std::is_same<T, float>::value) {
            // float32 output: the approximate square root of the float32-rounded variance is
            // within 2 ulp (1.2e-7) of rounding the double root; values outside the float range take the
            // double path (also negatives -> NaN, 0 -> 0)
            if (v > 1e-30 && v < 1e30) {
                float r;  // MUFU.SQRT: <= 1 ulp, no slow-path call (the _rn form carries one per output)
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)v));
                return r;
            }
            return (T)ieee_sqrt(v);
        }
        return (T)(SQRT ? fast_sqrt(v) : v);
    }
    // Branch-free form for the prefix kernel: the float32 root of the float32-rounded variance.  `suspect`
    // is raised when that image is not a normal finite number (zero, subnormal, inf: the double value may
    // be tiny or huge); the caller then redoes the tile's outputs with finalize_fast (one branch per tile).
    __device__ static __forceinline__ T finalize_pfx(const double *s, double rc, double rc1, bool &suspect) {
        const double v = dmul(dsub(s[1], dmul(dmul(s[0], s[0]), rc)), rc1);
        if constexpr (SQRT && std::is_same<T, float>::value) {
            const float t = (float)v;
            suspect |= ((__float_as_uint(t) & 0x7fffffffu) - 0x00800000u) >= 0x7f000000u;
            float r;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
            return r;
        }
        return (T)(SQRT ? fast_sqrt(v) : v);
    }
};
template <typename T>
struct OpCov {
    static constexpr int NIN = 2, NCH = 3;
    static constexpr bool ADD_FIRST = false;
    static constexpr int MIN_COUNT_FLOOR = 2;  // moving.py:193
    __device__ static __forceinline__ void contrib(T a, T b, double *c) {
        c[0] = (double)a;
        c[1] = (double)b;
        c[2] = prod_as_input(a, b);
    }
    __device__ static __forceinline__ T finalize(const double *s, int count) {
        // (prodsum - asum * bsum / count) / (count - 1)   moving.py:218
        return (T)(dsub(s[2], dmul(s[0], s[1]) / (double)count) / (double)(count - 1));
    }
    __device__ static __forceinline__ T finalize_fast(const double *s, double rc, double rc1) {
        return (T)dmul(dsub(s[2], dmul(dmul(s[0], s[1]), rc)), rc1);
    }
    __device__ static __forceinline__ T finalize_pfx(const double *s, double rc, double rc1, bool &) { return finalize_fast(s, rc, rc1); }
};
template <typename T>
struct OpCorr {
    static constexpr int NIN = 2, NCH = 5;
    static constexpr bool ADD_FIRST = false;
    static constexpr int MIN_COUNT_FLOOR = 1;  // moving.py:232
    __device__ static __forceinline__ void contrib(T a, T b, double *c) {
        c[0] = (double)a;
        c[1] = (double)b;
        c[2] = prod_as_input(a, b);
        c[3] = prod_as_input(a, a);
        c[4] = prod_as_input(b, b);
    }
    __device__ static __forceinline__ T finalize(const double *s, int count) {
        // moving.py:262-272: population moments through count_reciprocal
        double rc = 1.0 / (double)count;
        double avg_a = dmul(s[0], rc), avg_b = dmul(s[1], rc);
        double var_a = dsub(dmul(s[3], rc), dmul(avg_a, avg_a));
        double var_b = dsub(dmul(s[4], rc), dmul(avg_b, avg_b));
        double cov = dsub(dmul(s[2], rc), dmul(avg_a, avg_b));
        double vv = dmul(var_a, var_b);
        return vv > 0 ? (T)(cov / sqrt(vv)) : quiet_nan<T>();
    }
    __device__ static __forceinline__ T finalize_fast(const double *s, double rc, double) {
        const double avg_a = dmul(s[0], rc), avg_b = dmul(s[1], rc);
        const double var_a = dsub(dmul(s[3], rc), dmul(avg_a, avg_a));
        const double var_b = dsub(dmul(s[4], rc), dmul(avg_b, avg_b));
        const double cov = dsub(dmul(s[2], rc), dmul(avg_a, avg_b));
        const double vv = dmul(var_a, var_b);
        if constexpr (std::is_same<T, float>::value) {
            if (vv > 1e-30 && vv < 1e30) return __fmul_rn((float)cov, rsqrtf((float)vv));  // ~2e-7 relative
        }
        if (!(vv > 0)) return quiet_nan<T>();
        if (vv > 1e-35 && vv < 1e35) return (T)dmul(cov, fast_rsqrt(vv));
        return (T)dmul(cov, ieee_rsqrt(vv));
    }
    __device__ static __forceinline__ T finalize_pfx(const double *s, double rc, double rc1, bool &suspect) {
        if constexpr (std::is_same<T, float>::value) {
            const double avg_a = dmul(s[0], rc), avg_b = dmul(s[1], rc);
            const double var_a = dsub(dmul(s[3], rc), dmul(avg_a, avg_a));
            const double var_b = dsub(dmul(s[4], rc), dmul(avg_b, avg_b));
            const double cov = dsub(dmul(s[2], rc), dmul(avg_a, avg_b));
            const float t = (float)dmul(var_a, var_b);
            // not a normal positive float (<= 0, subnormal, inf, NaN): exact path decides (NaN gate, tiny products)
            suspect |= (__float_as_uint(t) - 0x00800000u) >= 0x7f000000u;
            return __fmul_rn((float)cov, rsqrtf(t));
        }
        return finalize_fast(s, rc, rc1);
    }
};

template <class Op, typename T>
__device__ __forceinline__ bool obs_valid(T a, T b) {
    if (Op::NIN == 2) return !(is_nan(a) || is_nan(b));
    return !is_nan(a);
}

// Invalid observations contribute +0.0 to every channel and 0 to the count: one select on the
// INPUT value instead of a select per 64-bit channel (x + 0.0 == x for every running sum that
// can occur here; the sums start at +0.0).
template <class Op, typename T>
__device__ __forceinline__ void acc_add(double *s, int &count, T a, T b) {
    const bool v = obs_valid<Op>(a, b);
    double c[Op::NCH];
    Op::contrib(v ? a : (T)0, v ? b : (T)0, c);
#pragma unroll
    for (int q = 0; q < Op::NCH; q++) s[q] = dadd(s[q], c[q]);
    count += v ? 1 : 0;
}
template <class Op, typename T>
__device__ __forceinline__ void acc_sub(double *s, int &count, T a, T b) {
    const bool v = obs_valid<Op>(a, b);
    double c[Op::NCH];
    Op::contrib(v ? a : (T)0, v ? b : (T)0, c);
#pragma unroll
    for (int q = 0; q < Op::NCH; q++) s[q] = dsub(s[q], c[q]);
    count -= v ? 1 : 0;
}

// --------------------------------------------------------------------------- row-tile kernel
struct MoveParams {
    const void *a, *b;
    void *out;
    const void *a_halo, *b_halo;
    int64_t halo_len;
    int64_t rows, n;
    int window;     // halo must fit in shared memory on this path
    int min_count;  // already clamped per op, saturated to int
    int tiles_per_row;
    int prefetch_dist;  // tiles ahead whose spans this CTA prefetches into L2 (0: off)
    int64_t ntiles;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory carve-up (bytes), shared by host sizing and the kernel
template <typename T, int NIN, int NCH, int THREADS, int E, bool RCP>
struct MoveSmem {
    static constexpr int TILE = THREADS * E;
    static constexpr int NCHP = NCH + 1;  // channels + count
    // mbarrier + per-warp scan totals (2 sets of NCHP doubles per warp: delta scan, halo sum)
    __host__ __device__ static size_t header_bytes() { return 64 + (size_t)2 * NCHP * (THREADS / 32 + 1) * sizeof(double); }
    __host__ __device__ static size_t rcp_bytes(int window) { return RCP ? align16((size_t)(window + 2) * sizeof(double)) : 0; }
    __host__ __device__ static size_t in_bytes(int window) { return align16((size_t)(window + TILE) * sizeof(T) + 16); }
    __host__ __device__ static size_t out_bytes() { return align16((size_t)TILE * sizeof(T) + 16); }
    __host__ __device__ static size_t work_bytes(int) { return out_bytes(); }
    __host__ __device__ static size_t total(int window) {
        return header_bytes() + rcp_bytes(window) + NIN * in_bytes(window) + work_bytes(window);
    }
};

// Window states at every thread's chunk start from per-thread quantities:
//   delta[c] = (sum over the thread's E entering elements) - (sum over its E leaving ones)
//   halo[c]  = the thread's share of the `window` elements preceding the tile
// S(thread t) = sum_all(halo) + sum_{t' < t} delta(t')   -- one shuffle scan + one barrier.
// Every partial sum is bounded by a window sum, so the rounding error stays at a few ulps
// of the window sum (better than differencing tile-long prefixes).
template <int THREADS, int NCHP>
__device__ __forceinline__ void block_delta_scan(const double *delta, const double *halo, double *state,
                                                 double *scratch /* 2*NCHP*(THREADS/32+1) */) {
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double inc[NCHP], hs[NCHP];
#pragma unroll
    for (int c = 0; c < NCHP; c++) {
        inc[c] = delta[c];
        hs[c] = halo[c];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < NCHP; c++) {
            const double o = __shfl_up_sync(0xffffffffu, inc[c], d);
            if (lane >= d) inc[c] += o;
            hs[c] += __shfl_xor_sync(0xffffffffu, hs[c], d);
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < NCHP; c++) {
            scratch[c * NW + wid] = inc[c];
            scratch[(NCHP + c) * NW + wid] = hs[c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NCHP; c++) {
        double base = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) {
            base += scratch[(NCHP + c) * NW + w2];
            if (w2 < wid) base += scratch[c * NW + w2];
        }
        state[c] = base + (inc[c] - delta[c]);
    }
}

template <typename T, class Op, int THREADS, int E, bool RCP>
__global__ void __launch_bounds__(THREADS) move_rowtile_kernel(MoveParams p) {
    constexpr int NIN = Op::NIN, NCH = Op::NCH, NCHP = NCH + 1;
    constexpr int TILE = THREADS * E;
    using SM = MoveSmem<T, NIN, NCH, THREADS, E, RCP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x / p.tiles_per_row;
    const int tile = blockIdx.x % p.tiles_per_row;
    const int64_t c0 = (int64_t)tile * TILE;
    const int w = p.window;
    const int len = w + TILE;  // span length
    const int64_t p0 = c0 - w;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw + 64);
    double *rcp = reinterpret_cast<double *>(smem_raw + SM::header_bytes());  // rcp[c + 1] = 1 / c
    unsigned char *in_base = smem_raw + SM::header_bytes() + SM::rcp_bytes(w);
    unsigned char *work_base = in_base + NIN * SM::in_bytes(w);

    const T *row_a = reinterpret_cast<const T *>(p.a) + row * p.n;
    const T *row_b = NIN == 2 ? reinterpret_cast<const T *>(p.b) + row * p.n : nullptr;
    T *row_out = reinterpret_cast<T *>(p.out) + row * p.n;
    const T *halo_a = p.a_halo ? reinterpret_cast<const T *>(p.a_halo) + row * p.halo_len : nullptr;
    const T *halo_b = (NIN == 2 && p.b_halo) ? reinterpret_cast<const T *>(p.b_halo) + row * p.halo_len : nullptr;

    T *sa = reinterpret_cast<T *>(in_base + span_phase(row_a, p0));
    T *sb = NIN == 2 ? reinterpret_cast<T *>(in_base + SM::in_bytes(w) + span_phase(row_b, p0)) : nullptr;

    // ---- stage the span(s): one bulk copy each + thread-filled edges
    const SpanPlan<T> pla = span_plan(row_a, p0, len, p.n);
    SpanPlan<T> plb = pla;
    if (NIN == 2) plb = span_plan(row_b, p0, len, p.n);
    const uint32_t tx = pla.blk_bytes + (NIN == 2 ? plb.blk_bytes : 0u);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && tx > 0) {
        mbar_arrive_expect_tx(bar, tx);
        if (pla.blk_bytes) bulk_g2s(sa + pla.blk_lo, row_a + p0 + pla.blk_lo, pla.blk_bytes, bar);
        if (NIN == 2 && plb.blk_bytes) bulk_g2s(sb + plb.blk_lo, row_b + p0 + plb.blk_lo, plb.blk_bytes, bar);
    }
    if (tid == 0 && p.prefetch_dist > 0) {
        const int64_t tl = (int64_t)blockIdx.x + p.prefetch_dist;
        if (tl < p.ntiles) {
            const int64_t row2 = tl / p.tiles_per_row;
            const int64_t c2 = (tl % p.tiles_per_row) * (int64_t)TILE;
            // the tile proper; its halo is the tail of the preceding tile (fetched by that one)
            span_prefetch_l2(reinterpret_cast<const T *>(p.a) + row2 * p.n, c2, TILE, p.n);
            if (NIN == 2) span_prefetch_l2(reinterpret_cast<const T *>(p.b) + row2 * p.n, c2, TILE, p.n);
        }
    }
    span_fill_edges<T, THREADS>(sa, row_a, p0, len, pla, quiet_nan<T>(), halo_a, p.halo_len);
    if (NIN == 2) span_fill_edges<T, THREADS>(sb, row_b, p0, len, plb, quiet_nan<T>(), halo_b, p.halo_len);
    if (RCP) {
        // reciprocal table while the copy is in flight: rcp[c + 1] = 1.0 / c (IEEE division),
        // rcp[0] = rcp[1] = 0 so that count == 0 and count - 1 == -1 index harmless entries
        for (int c = tid; c <= w + 1; c += THREADS) rcp[c] = c >= 2 ? 1.0 / (double)(c - 1) : 0.0;
    }
    if (tx > 0) mbar_wait(bar, 0);
    __syncthreads();

    // ---- window state at this thread's chunk start: observations at j in [t*E, t*E + w)
    double s[NCH];
#pragma unroll
    for (int q = 0; q < NCH; q++) s[q] = 0.0;
    int count = 0;
    const int jt = tid * E;  // trailing index of this thread's first output
    if (w <= kDirectMax) {
        for (int j = jt; j < jt + w; j++) acc_add<Op, T>(s, count, sa[j], NIN == 2 ? sb[j] : sa[j]);
    } else {
        // pass 1: this thread's entering-minus-leaving totals + its share of the halo sum
        double delta[NCHP], halo[NCHP], st[NCHP];
        {
            double cs[NCH];
#pragma unroll
            for (int q = 0; q < NCH; q++) cs[q] = 0.0;
            int cc = 0;
#pragma unroll
            for (int k = 0; k < E; k++) {
                acc_add<Op, T>(cs, cc, sa[jt + w + k], NIN == 2 ? sb[jt + w + k] : sa[jt + w + k]);
                acc_sub<Op, T>(cs, cc, sa[jt + k], NIN == 2 ? sb[jt + k] : sa[jt + k]);
            }
#pragma unroll
            for (int q = 0; q < NCH; q++) delta[q] = cs[q];
            delta[NCH] = (double)cc;
#pragma unroll
            for (int q = 0; q < NCH; q++) cs[q] = 0.0;
            cc = 0;
            for (int j = tid; j < w; j += THREADS) acc_add<Op, T>(cs, cc, sa[j], NIN == 2 ? sb[j] : sa[j]);
#pragma unroll
            for (int q = 0; q < NCH; q++) halo[q] = cs[q];
            halo[NCH] = (double)cc;
        }
        block_delta_scan<THREADS, NCHP>(delta, halo, st, scratch);
#pragma unroll
        for (int q = 0; q < NCH; q++) s[q] = st[q];
        count = __double2int_rn(st[NCH]);
    }

    // ---- running window over this thread's E outputs
    T *sout = reinterpret_cast<T *>(work_base + span_phase(row_out, c0));
    const int mc = p.min_count;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const int jl = jt + w + k;  // entering element
        const int jr = jt + k;      // leaving element (NaN-filled before the row start)
        const T al = sa[jl], ar = sa[jr];
        const T bl = NIN == 2 ? sb[jl] : al, br = NIN == 2 ? sb[jr] : ar;
        if (Op::ADD_FIRST) {
            acc_add<Op, T>(s, count, al, bl);
            acc_sub<Op, T>(s, count, ar, br);
        } else {
            acc_sub<Op, T>(s, count, ar, br);
            acc_add<Op, T>(s, count, al, bl);
        }
        T r;
        if (RCP) {
            r = Op::finalize_fast(s, rcp[count + 1], rcp[count]);
        } else {
            r = Op::finalize(s, count);
        }
        sout[jt + k] = (count >= mc) ? r : quiet_nan<T>();
    }

    // ---- drain the out tile
    fence_async_smem();
    __syncthreads();
    const int64_t rem = p.n - c0;
    const int cnt = rem < TILE ? (int)rem : TILE;
    span_store<T, THREADS>(row_out + c0, sout, cnt);
    if (tid == 0) bulk_wait_read_all();
}

}  // namespace nbg
#include "nbg_move_prefix.cuh"
namespace nbg {

// ------------------------------------------------------------------------ column-walk kernel
// (outer, n, inner) C-contiguous, inner > 1.  Thread <-> (outer, segment, inner column).
struct MoveColParams {
    const void *a, *b;
    void *out;
    const void *a_halo, *b_halo;
    int64_t halo_len;
    int64_t outer, n, inner;
    int64_t window;
    int64_t min_count;
    int64_t seg_len;  // core positions per thread
    int64_t nseg;
};

template <typename T, class Op>
__global__ void __launch_bounds__(256) move_colwalk_kernel(MoveColParams p) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = p.outer * p.nseg * p.inner;
    if (gid >= total) return;
    const int64_t col = gid % p.inner;
    const int64_t seg = (gid / p.inner) % p.nseg;
    const int64_t o = gid / (p.inner * p.nseg);
    constexpr int NIN = Op::NIN, NCH = Op::NCH;
    const T *a = reinterpret_cast<const T *>(p.a) + o * p.n * p.inner + col;
    const T *b = NIN == 2 ? reinterpret_cast<const T *>(p.b) + o * p.n * p.inner + col : a;
    const T *ha = p.a_halo ? reinterpret_cast<const T *>(p.a_halo) + o * p.halo_len * p.inner + col : nullptr;
    const T *hb = (NIN == 2 && p.b_halo) ? reinterpret_cast<const T *>(p.b_halo) + o * p.halo_len * p.inner + col : ha;
    T *out = reinterpret_cast<T *>(p.out) + o * p.n * p.inner + col;
    const int64_t st = p.inner;
    const int64_t i0 = seg * p.seg_len;
    const int64_t i1 = min(i0 + p.seg_len, p.n);
    const int64_t w = p.window;

    auto load = [&](const T *x, const T *h, int64_t i) -> T {
        if (i >= 0) return x[i * st];
        if (h != nullptr) {
            int64_t hi = p.halo_len + i;
            if (hi >= 0) return h[hi * st];
        }
        return quiet_nan<T>();
    };

    double s[NCH];
#pragma unroll
    for (int q = 0; q < NCH; q++) s[q] = 0.0;
    int count = 0;
    // rebuild the window preceding i0 (empty for the first segment without halo)
    {
        int64_t lo = i0 - w;
        if (lo < -p.halo_len) lo = -p.halo_len;
        if (ha == nullptr && lo < 0) lo = 0;
        for (int64_t i = lo; i < i0; i++) acc_add<Op, T>(s, count, load(a, ha, i), NIN == 2 ? load(b, hb, i) : (T)0);
    }
    const int mc = (int)min(p.min_count, (int64_t)INT32_MAX);
    for (int64_t i = i0; i < i1; i++) {
        const T al = a[i * st];
        const T bl = NIN == 2 ? b[i * st] : al;
        const T ar = load(a, ha, i - w);
        const T br = NIN == 2 ? load(b, hb, i - w) : ar;
        if (Op::ADD_FIRST) {
            acc_add<Op, T>(s, count, al, bl);
            acc_sub<Op, T>(s, count, ar, br);
        } else {
            acc_sub<Op, T>(s, count, ar, br);
            acc_add<Op, T>(s, count, al, bl);
        }
        out[i * st] = (count >= mc) ? Op::finalize(s, count) : quiet_nan<T>();
    }
}

// ---------------------------------------------------------------------------------- launch
template <typename T>
struct TileCfg;
template <>
struct TileCfg<float> {
    static constexpr int THREADS = 512, E = 17;
};
template <>
struct TileCfg<double> {
    static constexpr int THREADS = 512, E = 9;
};

// Small problems (fewer than ~2 waves of the 512-thread tiles: BASELINE config 1 is 300 of them on 296
// slots, i.e. one full wave and a 4-CTA tail) take 128-thread tiles instead: four times as many CTAs, all
// resident at once, each with a shorter load -> compute -> store chain.
template <typename T, class Op, bool RCP, int THREADS = TileCfg<T>::THREADS>
static int launch_rowtile(MoveParams p, int64_t outer, int64_t n, cudaStream_t stream) {
    constexpr int E = TileCfg<T>::E;
    using SM = MoveSmem<T, Op::NIN, Op::NCH, THREADS, E, RCP>;
    if constexpr (THREADS == TileCfg<T>::THREADS) {
        const int64_t big_tiles = ((n + SM::TILE - 1) / SM::TILE) * outer;
        if (big_tiles < 4 * (int64_t)kNumSMs && !getenv("NBG_MOVE_BIG_TILES")) return launch_rowtile<T, Op, RCP, 128>(p, outer, n, stream);
    }
    const int64_t tpr = (n + SM::TILE - 1) / SM::TILE;
    if (tpr * outer > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_move: more than 2^31 tiles");
    p.tiles_per_row = (int)tpr;
    p.ntiles = tpr * outer;
    p.prefetch_dist = prefetch_distance(2, (size_t)Op::NIN * SM::TILE * sizeof(T));
    auto kern = move_rowtile_kernel<T, Op, THREADS, E, RCP>;
    int rc = allow_big_smem(kern, "nbg_move: cudaFuncSetAttribute");
    if (rc) return rc;
    kern<<<(unsigned)(tpr * outer), THREADS, SM::total(p.window), stream>>>(p);
    return check_launch("nbg_move(rowtile)");
}

template <typename T, class Op>
static int launch_move(const void *a, const void *b, void *out, int64_t outer, int64_t n, int64_t inner,
                       int64_t window, int64_t min_count, const void *a_halo, const void *b_halo,
                       int64_t halo_len, cudaStream_t stream) {
    if (min_count < Op::MIN_COUNT_FLOOR) min_count = Op::MIN_COUNT_FLOOR;
    if (outer * n * inner == 0) return NBG_OK;
    constexpr int THREADS = TileCfg<T>::THREADS, E = TileCfg<T>::E;
    using SMfast = MoveSmem<T, Op::NIN, Op::NCH, THREADS, E, true>;
    using SMslow = MoveSmem<T, Op::NIN, Op::NCH, THREADS, E, false>;
    if constexpr (std::is_same<T, float>::value) {
        // float32 data, wide windows: prefix differences (every observation widened once)
        // Measured on config 4 (float32, window 1000; running-window kernel -> prefix kernel): move_std 3.91 -> 3.74 ms,
        // move_cov 6.94 -> 5.91 ms; move_var 3.56 -> 3.58, move_mean 2.50 -> 2.64, move_corr 10.06 -> 10.37 ms.  Only
        // the ops it wins on take it (NBG_PFX=all / off force it on / off for A-B runs).
        constexpr bool kWins = std::is_same<Op, OpVar<T, true>>::value || std::is_same<Op, OpCov<T>>::value;
        const char *pe = getenv("NBG_PFX");
        const bool want = pe ? (strcmp(pe, "all") == 0 || (strcmp(pe, "off") != 0 && kWins)) : kWins;
        if (inner == 1 && window > kDirectMax && want && prefix_fits<T, Op>(window)) {
            MovePfxParams q = {};
            q.a = a, q.b = b, q.out = out, q.a_halo = a_halo, q.b_halo = b_halo, q.halo_len = halo_len;
            q.rows = outer, q.n = n, q.window = (int)window;
            q.min_count = (int)(min_count > INT32_MAX ? INT32_MAX : min_count);
            return launch_prefix<T, Op>(q, outer, n, stream);
        }
    }
    if (inner == 1 && window <= (1 << 20) && SMslow::total((int)window) <= kMaxSmem) {
        MoveParams p;
        p.a = a, p.b = b, p.out = out, p.a_halo = a_halo, p.b_halo = b_halo, p.halo_len = halo_len;
        p.rows = outer, p.n = n, p.window = (int)window;
        p.min_count = (int)(min_count > INT32_MAX ? INT32_MAX : min_count);
        p.tiles_per_row = 0;
        if (window <= kRcpMax && SMfast::total((int)window) <= kMaxSmem)
            return launch_rowtile<T, Op, true>(p, outer, n, stream);
        return launch_rowtile<T, Op, false>(p, outer, n, stream);
    }
    // column walk (also the fallback for windows whose halo does not fit in shared memory)
    MoveColParams p;
    p.a = a, p.b = b, p.out = out, p.a_halo = a_halo, p.b_halo = b_halo, p.halo_len = halo_len;
    p.outer = outer, p.n = n, p.inner = inner, p.window = window, p.min_count = min_count;
    // enough threads to fill the machine, but segments long enough to amortise the
    // window rebuild (window extra reads per segment)
    const int64_t cols = outer * inner;
    const int64_t want_threads = (int64_t)kNumSMs * 2048;
    int64_t nseg = (want_threads + cols - 1) / cols;
    int64_t min_seg = window * 8 > 64 ? window * 8 : 64;
    int64_t max_nseg = (n + min_seg - 1) / min_seg;
    if (nseg > max_nseg) nseg = max_nseg;
    if (nseg < 1) nseg = 1;
    p.seg_len = (n + nseg - 1) / nseg;
    p.nseg = (n + p.seg_len - 1) / p.seg_len;
    const int64_t total = cols * p.nseg;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_move: grid too large");
    move_colwalk_kernel<T, Op><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return check_launch("nbg_move(colwalk)");
}

template <typename T>
static int dispatch_move(int op, const void *a, const void *b, void *out, int64_t outer, int64_t n,
                         int64_t inner, int64_t window, int64_t min_count, const void *a_halo,
                         const void *b_halo, int64_t halo_len, cudaStream_t stream) {
#define NBG_MOVE_CASE(OPC, OP) \
    case OPC:                  \
        return launch_move<T, OP>(a, b, out, outer, n, inner, window, min_count, a_halo, b_halo, halo_len, stream)
    using OpStd = OpVar<T, true>;
    using OpVariance = OpVar<T, false>;
    switch (op) {
        NBG_MOVE_CASE(NBG_MOVE_MEAN, OpMean<T>);
        NBG_MOVE_CASE(NBG_MOVE_SUM, OpSum<T>);
        NBG_MOVE_CASE(NBG_MOVE_STD, OpStd);
        NBG_MOVE_CASE(NBG_MOVE_VAR, OpVariance);
        NBG_MOVE_CASE(NBG_MOVE_COV, OpCov<T>);
        NBG_MOVE_CASE(NBG_MOVE_CORR, OpCorr<T>);
        default:
            return fail(NBG_ERR_BAD_OP, "nbg_move: unknown op");
    }
#undef NBG_MOVE_CASE
}

}  // namespace nbg

extern "C" int nbg_move(int op, int dtype, const void *a, const void *b, void *out, int64_t outer, int64_t n,
                        int64_t inner, int64_t window, int64_t min_count, const void *a_halo, const void *b_halo,
                        int64_t halo_len, void *stream) {
    using namespace nbg;
    if (outer < 0 || n < 0 || inner < 0) return fail(NBG_ERR_BAD_ARG, "nbg_move: negative size");
    if (window <= 0) return fail(NBG_ERR_BAD_ARG, "nbg_move: window must be positive");
    if (min_count < 0) return fail(NBG_ERR_BAD_ARG, "nbg_move: min_count must be >= 0");
    if (halo_len < 0) return fail(NBG_ERR_BAD_ARG, "nbg_move: halo_len must be >= 0");
    const bool two = (op == NBG_MOVE_COV || op == NBG_MOVE_CORR);
    if (outer * n * inner > 0 && (!a || !out || (two && !b))) return fail(NBG_ERR_BAD_ARG, "nbg_move: null pointer");
    if (halo_len > 0 && (!a_halo || (two && !b_halo))) return fail(NBG_ERR_BAD_ARG, "nbg_move: null halo");
    if (halo_len == 0) a_halo = b_halo = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case NBG_F32:
            return dispatch_move<float>(op, a, b, out, outer, n, inner, window, min_count, a_halo, b_halo, halo_len, st);
        case NBG_F64:
            return dispatch_move<double>(op, a, b, out, outer, n, inner, window, min_count, a_halo, b_halo, halo_len, st);
        default:
            return fail(NBG_ERR_BAD_DTYPE, "nbg_move: dtype must be NBG_F32 or NBG_F64");
    }
}
