// nbg_reduce.cu -- plain NaN-aware reductions over the middle axis of an (outer, n, inner)
// array: allnan, anynan, nancount, nansum, nanmean, nanvar, nanstd, nanargmax, nanargmin,
// nanmax, nanmin (numbagg/funcs.py:23-242 behind ndaggregate / ndreduce,
// numbagg/decorators.py:188-260, 906-1031).
//
// All of it is one streaming pass at HBM speed.  Every op is a "reducer": a small state, an
// `add` for one element, an order-independent `merge`, and a `finalize`.  Three kernels walk
// the data, chosen by shape (red_geometry):
//   rows_cta   inner == 1, long rows: one CTA per (row, segment), 16-byte streaming loads;
//   group      short rows / few elements per output: G lanes per output (G = 1..32);
//   cols       inner > 1: the CTA reads a contiguous (segment rows x w columns) tile with
//              consecutive threads on consecutive addresses, so a thread always sees one
//              column; threads that share a column merge through shared memory.
// Segments exist only to fill the machine when there are few outputs; their states go to the
// workspace as 3-word records and a merge kernel folds them.  The same records are the
// exchange format for element-sharded (multi-GPU) reductions: nbg_reduce_partial /
// nbg_reduce_merge.
//
// Numerics: sums of floats accumulate in double (the reference accumulates nansum in the
// input dtype sequentially; ours is at least as accurate), variance is a batched two-pass in
// registers folded with Chan's pairwise update (robust like the reference's two loops, one
// read of the data).  Counts, flags, extrema and arg-extrema are exact.
#include "nbg_common.cuh"

#include <stdlib.h>

#include <type_traits>

namespace nbg {
namespace {

typedef unsigned long long u64;
typedef long long i64;

constexpr int kRedThreads = 256;
constexpr int kStateWords = NBG_REDUCE_STATE_WORDS;

// Where a kernel's result goes: the typed output (finalize) or 3-word state records
// states[(part * 3 + word) * outs + j] (a segment of this launch, or a caller's shard).
struct RedOut {
    void *out;
    u64 *states;
    i64 outs;
    i64 part;
    int emit_state;
    i64 n_total;
    i64 ddof;
};

__device__ __forceinline__ u64 d2w(double x) { return (u64)__double_as_longlong(x); }
__device__ __forceinline__ double w2d(u64 w) { return __longlong_as_double((i64)w); }

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// NaN -> 0 in the input type.  Written as `nan ? 0 : v` (even through bit casts) the
// compiler widens first and then selects both halves of the double: two FSELs per element.
__device__ __forceinline__ float nan_to_zero(float v) {
    float r;  // inline PTX keeps the select where it is written
    asm("{ .reg .pred p; setp.nan.f32 p, %1, %1; selp.f32 %0, 0f00000000, %1, p; }" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ double nan_to_zero(double v) { return v != v ? 0.0 : v; }
__device__ __forceinline__ int32_t nan_to_zero(int32_t v) { return v; }
__device__ __forceinline__ int64_t nan_to_zero(int64_t v) { return v; }

// 1 / c for a positive count c: hardware float reciprocal + two Newton steps in double
// (relative error ~1e-29 before the final rounding) instead of the ~25-instruction division.
__device__ __forceinline__ double fast_rcp(double c) {
    float y0;  // MUFU.RCP only: the correctly rounded __frcp_rn drags in a slow-path call
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)c));
    double y = (double)y0;
    double e = fma(-c, y, 1.0);
    y = fma(y, e, y);
    e = fma(-c, y, 1.0);
    return fma(y, e, y);
}

// NaN-ignoring extreme of B (a power of two) values as a balanced tree: NaN only if all are
__device__ __forceinline__ float ext2(float a, float b, bool mx) { return mx ? fmaxf(a, b) : fminf(a, b); }
__device__ __forceinline__ double ext2(double a, double b, bool mx) { return mx ? fmax(a, b) : fmin(a, b); }
template <bool MAX, typename K, int B>
__device__ __forceinline__ K tree_extreme(const K (&x)[B]) {
    static_assert((B & (B - 1)) == 0, "batch sizes are powers of two");
    K t[B];
#pragma unroll
    for (int b = 0; b < B; b++) t[b] = x[b];
#pragma unroll
    for (int w = B / 2; w >= 1; w /= 2)
#pragma unroll
        for (int b = 0; b < w; b++) t[b] = ext2(t[b], t[b + w], MAX);
    return t[0];
}

// ------------------------------------------------------------------------------- reducers
// MODE 0 allnan, 1 anynan, 2 nancount -- funcs.py:23-68.
template <typename T, int MODE>
struct RCount {
    using In = T;
    static constexpr bool TWO_PASS = false;
    static constexpr bool ORDERED = false;
    static constexpr int MIN_CTAS = 6;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = true;
    struct State {
        i64 c;
    };
    static __device__ __forceinline__ State init() { return {0}; }
    static __device__ __forceinline__ void add(State &s, T v, i64) { s.c += is_nan(v) ? 0 : 1; }
    template <int B, int V, bool FULL>
    static __device__ __forceinline__ void add_batch(State &s, const T (&v)[B], uint32_t mask, i64, i64) {
        int c = 0;  // 32-bit inside the batch, one 64-bit add per batch
#pragma unroll
        for (int b = 0; b < B; b++) c += ((FULL || ((mask >> b) & 1u)) && !is_nan(v[b])) ? 1 : 0;
        s.c += c;
    }
    static __device__ __forceinline__ void merge(State &a, const State &b) { a.c += b.c; }
    static __device__ __forceinline__ State shfl(const State &s, int m) { return {shfl_xor_t(s.c, m)}; }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        w[0] = (u64)s.c;
        w[1] = 0;
        w[2] = 0;
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) { return {(i64)w[0]}; }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        if (MODE == 0)
            ((uint8_t *)o.out)[j] = s.c == 0;
        else if (MODE == 1)
            ((uint8_t *)o.out)[j] = s.c < o.n_total;
        else
            ((i64 *)o.out)[j] = s.c;
    }
};

// nansum -- funcs.py:71-85.  Floats: double accumulator; ints: int64 (wraps like the int32
// loop after the final narrowing).
template <typename T>
struct RSum {
    using In = T;
    static constexpr bool TWO_PASS = false;
    static constexpr bool ORDERED = false;
    static constexpr int MIN_CTAS = 6;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = false;
    static constexpr bool IS_INT = std::is_integral<T>::value;
    using Acc = typename std::conditional<IS_INT, i64, double>::type;
    struct State {
        Acc s;
    };
    static __device__ __forceinline__ State init() { return {Acc(0)}; }
    static __device__ __forceinline__ void add(State &s, T v, i64) { s.s += (Acc)nan_to_zero(v); }
    static __device__ __forceinline__ void merge(State &a, const State &b) { a.s += b.s; }
    static __device__ __forceinline__ State shfl(const State &s, int m) { return {shfl_xor_t(s.s, m)}; }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        if constexpr (IS_INT)
            w[0] = (u64)s.s;
        else
            w[0] = d2w(s.s);
        w[1] = 0;
        w[2] = 0;
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) {
        if constexpr (IS_INT)
            return {(i64)w[0]};
        else
            return {w2d(w[0])};
    }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        ((T *)o.out)[j] = (T)s.s;
    }
};

// nanmean -- funcs.py:88-104 (double sum, int64 count, one division).
template <typename T>
struct RMean {
    using In = T;
    static constexpr bool TWO_PASS = false;
    static constexpr bool ORDERED = false;
    static constexpr int MIN_CTAS = 6;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = true;
    struct State {
        double s;
        i64 c;
    };
    static __device__ __forceinline__ State init() { return {0.0, 0}; }
    static __device__ __forceinline__ void add(State &s, T v, i64) {
        s.s += (double)nan_to_zero(v);
        s.c += is_nan(v) ? 0 : 1;
    }
    template <int B, int V, bool FULL>
    static __device__ __forceinline__ void add_batch(State &s, const T (&v)[B], uint32_t mask, i64, i64) {
        int c = 0;
#pragma unroll
        for (int b = 0; b < B; b++) {
            const bool in = FULL || ((mask >> b) & 1u);
            s.s += (double)nan_to_zero(in ? v[b] : T(0));
            c += (in && !is_nan(v[b])) ? 1 : 0;
        }
        s.c += c;
    }
    static __device__ __forceinline__ void merge(State &a, const State &b) {
        a.s += b.s;
        a.c += b.c;
    }
    static __device__ __forceinline__ State shfl(const State &s, int m) {
        return {shfl_xor_t(s.s, m), shfl_xor_t(s.c, m)};
    }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        w[0] = (u64)s.c;
        w[1] = d2w(s.s);
        w[2] = 0;
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) { return {w2d(w[1]), (i64)w[0]}; }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        ((T *)o.out)[j] = s.c > 0 ? (T)(s.s / (double)s.c) : quiet_nan<T>();
    }
};

// nanvar / nanstd -- funcs.py:107-158.  The reference makes two passes (mean, then squared
// deviations).  We read once: each register batch is reduced two-pass (batch mean, then
// centred squares), and batches / threads / segments fold with Chan's update
//   mean = mean_a + d * cb / c,  M2 = M2_a + M2_b + d^2 * ca * cb / c,   d = mean_b - mean_a
// whose terms are all centred, so no cancellation arises wherever the data sits.
template <typename T, bool SQRT>
struct RVar {
    using In = T;
    static constexpr bool TWO_PASS = true;  // rows_tile: mean first, then centred squares
    static constexpr bool ORDERED = false;
    static constexpr int MIN_CTAS = 4;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = true;
    struct State {
        i64 c;
        double mean, m2;
    };
    static __device__ __forceinline__ State init() { return {0, 0.0, 0.0}; }
    static __device__ __forceinline__ void merge(State &a, const State &b) {
        if (b.c == 0) return;
        if (a.c == 0) {
            a = b;
            return;
        }
        const double ca = (double)a.c, cb = (double)b.c;
        const double r = cb * fast_rcp(ca + cb);
        const double d = b.mean - a.mean;
        a.mean = fma(d, r, a.mean);
        a.m2 = a.m2 + b.m2 + d * d * ca * r;
        a.c += b.c;
    }
    static __device__ __forceinline__ void add(State &s, T v, i64) {
        if (!is_nan(v)) merge(s, State{1, (double)v, 0.0});
    }
    // One register batch, two passes.  Pass 1 runs in the INPUT type: all it must deliver is
    // a centre K near the batch mean (any K works; the closer, the smaller s1).  Missing
    // elements are replaced by K itself, so pass 2 needs no masks: their deviation is 0.
    template <int B, int V, bool FULL>
    static __device__ __forceinline__ void add_batch(State &s, const T (&v)[B], uint32_t mask, i64, i64) {
        uint32_t okm = 0;
        T sum0 = T(0), sum1 = T(0);  // two chains per pass: half the dependent latency
#pragma unroll
        for (int b = 0; b < B; b++) {
            const bool ok = (FULL || ((mask >> b) & 1u)) && !is_nan(v[b]);
            okm |= (ok ? 1u : 0u) << b;
            if (b & 1)
                sum1 += ok ? v[b] : T(0);
            else
                sum0 += ok ? v[b] : T(0);
        }
        const int cb = __popc(okm);
        if (cb == 0) return;
        const double rb = fast_rcp((double)cb);
        T k = (sum0 + sum1) * (T)rb;
        // overflowed / infinite sum: any finite centre will do
        if constexpr (sizeof(T) == 4) {
            if (!(fabsf((float)k) <= 3.0e38f)) k = T(0);
        } else {
            if (!(fabs((double)k) <= 1.7e308)) k = T(0);
        }
        const double kd = (double)k;
        double s1, s2;
        if constexpr (sizeof(T) == 4) {
            // float32 data: deviations from the batch centre in float32 (x - k is exact
            // whenever x and k are within a factor of two -- Sterbenz -- and rounds at 6e-8
            // otherwise); 16 of them are summed in float32 before joining the double state.
            // Relative error of the variance <~ 1e-6, inside the float32 result's tolerance.
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
            for (int b = 0; b < B; b++) {
                const float d = (((okm >> b) & 1u) ? v[b] : k) - k;
                if (b & 1) {
                    s1b += d;
                    s2b = fmaf(d, d, s2b);
                } else {
                    s1a += d;
                    s2a = fmaf(d, d, s2a);
                }
            }
            s1 = (double)(s1a + s1b);
            s2 = (double)s2a + (double)s2b;
        } else {
            double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
#pragma unroll
            for (int b = 0; b < B; b++) {
                const T x = ((okm >> b) & 1u) ? v[b] : k;
                const double d = (double)x - kd;
                if (b & 1) {
                    s1b += d;
                    s2b = fma(d, d, s2b);
                } else {
                    s1a += d;
                    s2a = fma(d, d, s2a);
                }
            }
            s1 = s1a + s1b;
            s2 = s2a + s2b;
        }
        State bs;
        bs.c = cb;
        bs.mean = fma(s1, rb, kd);
        bs.m2 = s2 - s1 * s1 * rb;
        if (bs.m2 < 0.0) bs.m2 = 0.0;  // rounding only; NaN (inf - inf, as in the reference) stays NaN
        merge(s, bs);
    }
    static __device__ __forceinline__ State shfl(const State &s, int m) {
        return {shfl_xor_t(s.c, m), shfl_xor_t(s.mean, m), shfl_xor_t(s.m2, m)};
    }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        w[0] = (u64)s.c;
        w[1] = d2w(s.mean);
        w[2] = d2w(s.m2);
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) { return {(i64)w[0], w2d(w[1]), w2d(w[2])}; }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        double r = quiet_nan<double>();
        if (s.c > o.ddof) {
            r = s.m2 / (double)(s.c - o.ddof);
            if (SQRT) r = sqrt(r);
        }
        ((T *)o.out)[j] = (T)r;
    }
};

// nanmax / nanmin -- funcs.py:200-242: `if ai >= amax` from -inf, NaN when nothing passed.
// Floats stay in their type; ints compare as ints (the reference compares them as doubles,
// a monotone map) and come back as int64 through the same double round trip.
template <typename T, bool MAX>
struct RExt {
    using In = T;
    static constexpr bool TWO_PASS = false;
    static constexpr bool ORDERED = false;
    static constexpr int MIN_CTAS = 6;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = true;
    static constexpr bool IS_INT = std::is_integral<T>::value;
    struct State {
        T m;
        int any;
    };
    static __device__ __forceinline__ T lowest() {
        if constexpr (IS_INT)
            return MAX ? (sizeof(T) == 4 ? (T)INT32_MIN : (T)INT64_MIN) : (sizeof(T) == 4 ? (T)INT32_MAX : (T)INT64_MAX);
        else
            return MAX ? (T)-INFINITY : (T)INFINITY;
    }
    static __device__ __forceinline__ State init() { return {lowest(), 0}; }
    static __device__ __forceinline__ void add(State &s, T v, i64) {
        const bool ok = MAX ? v >= s.m : v <= s.m;
        s.m = ok ? v : s.m;
        s.any |= ok ? 1 : 0;
    }
    // a batch folds to its own extreme first (one min/max per element, NaN ignored by the
    // hardware min/max), then meets the running value once
    template <int B, int V, bool FULL>
    static __device__ __forceinline__ void add_batch(State &s, const T (&v)[B], uint32_t mask, i64, i64) {
        if constexpr (IS_INT) {
            T m = lowest();
#pragma unroll
            for (int b = 0; b < B; b++) {
                const T x = (FULL || ((mask >> b) & 1u)) ? v[b] : lowest();
                m = MAX ? (x > m ? x : m) : (x < m ? x : m);
            }
            s.m = MAX ? (m > s.m ? m : s.m) : (m < s.m ? m : s.m);
            s.any |= (FULL || mask != 0) ? 1 : 0;
        } else {
            T x[B];
#pragma unroll
            for (int b = 0; b < B; b++) x[b] = (FULL || ((mask >> b) & 1u)) ? v[b] : quiet_nan<T>();
            const T m = tree_extreme<MAX>(x);
            s.m = ext2(s.m, m, MAX);
            s.any |= m == m ? 1 : 0;
        }
    }
    static __device__ __forceinline__ void merge(State &a, const State &b) {
        if (b.any) add(a, b.m, 0);
    }
    static __device__ __forceinline__ State shfl(const State &s, int m) {
        return {shfl_xor_t(s.m, m), shfl_xor_t(s.any, m)};
    }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        w[0] = (u64)s.any;
        if constexpr (IS_INT)
            w[1] = (u64)(i64)s.m;
        else
            w[1] = d2w((double)s.m);
        w[2] = 0;
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) {
        if constexpr (IS_INT)
            return {(T)(i64)w[1], (int)w[0]};
        else
            return {(T)w2d(w[1]), (int)w[0]};
    }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        if constexpr (IS_INT) {
            // int64(float64(amax)); an empty slice is rejected by the caller beforehand
            ((i64 *)o.out)[j] = s.any ? __double2ll_rz((double)s.m) : (i64)0;
        } else {
            ((T *)o.out)[j] = s.any ? s.m : quiet_nan<T>();
        }
    }
};

// nanargmax / nanargmin -- funcs.py:161-197: strict compare from -inf, or the first non-NaN
// while nothing was taken; so the first occurrence of the extreme wins.  Ints are compared
// as doubles, exactly as numba types `ai > amax` with amax = -np.inf.  idx = -1: nothing
// taken (the caller raises ValueError like the reference).
template <typename T, bool MAX>
struct RArg {
    using In = T;
    static constexpr bool TWO_PASS = false;
    static constexpr bool ORDERED = true;  // add() needs increasing indices per thread
    static constexpr int MIN_CTAS = 5;  // resident 256-thread CTAs per SM the kernels are compiled for
    static constexpr bool BATCHED = true;
    static constexpr bool IS_INT = std::is_integral<T>::value;
    using K = typename std::conditional<IS_INT, double, T>::type;
    struct State {
        K key;
        i64 idx;
    };
    static __device__ __forceinline__ State init() { return {MAX ? (K)-INFINITY : (K)INFINITY, -1}; }
    static __device__ __forceinline__ bool better(K a, K b) { return MAX ? a > b : a < b; }
    // elements reach one thread in increasing index order
    static __device__ __forceinline__ void add(State &s, T v, i64 idx) {
        const K k = (K)v;
        const bool take = better(k, s.key) || (s.idx < 0 && !is_nan(k));
        s.key = take ? k : s.key;
        s.idx = take ? idx : s.idx;
    }
    // The running extreme changes O(log n) times per thread, so a batch is folded to its
    // extreme with one min/max per element and only a batch that beats the running value
    // (or is the first with data) looks for the position of its first extreme.
    template <int B, int V, bool FULL>
    static __device__ __forceinline__ void add_batch(State &s, const T (&v)[B], uint32_t mask, i64 idx0, i64 kstride) {
        K x[B];
#pragma unroll
        for (int b = 0; b < B; b++) x[b] = (FULL || ((mask >> b) & 1u)) ? (K)v[b] : quiet_nan<K>();
        const K m = tree_extreme<MAX>(x);
        if (better(m, s.key) || (s.idx < 0 && m == m)) {
            int first = 0;
#pragma unroll
            for (int b = B - 1; b >= 0; b--) first = x[b] == m ? b : first;
            s.key = m;
            s.idx = idx0 + (i64)(first / V) * kstride + (first % V);
        }
    }
    static __device__ __forceinline__ void merge(State &a, const State &b) {
        const bool take =
            b.idx >= 0 && (a.idx < 0 || better(b.key, a.key) || (b.key == a.key && b.idx < a.idx));
        a.key = take ? b.key : a.key;
        a.idx = take ? b.idx : a.idx;
    }
    static __device__ __forceinline__ State shfl(const State &s, int m) {
        return {shfl_xor_t(s.key, m), shfl_xor_t(s.idx, m)};
    }
    static __device__ __forceinline__ void pack(const State &s, u64 (&w)[3]) {
        w[0] = (u64)s.idx;
        w[1] = d2w((double)s.key);
        w[2] = 0;
    }
    static __device__ __forceinline__ State unpack(const u64 (&w)[3]) { return {(K)w2d(w[1]), (i64)w[0]}; }
    static __device__ __forceinline__ void finalize(const State &s, const RedOut &o, i64 j) {
        ((i64 *)o.out)[j] = s.idx;
    }
};

// B elements at once.  Element b has index idx0 + (b / V) * kstride + (b % V); `mask` (used
// when !FULL) says which elements exist.
template <typename R, int B, int V, bool FULL>
__device__ __forceinline__ void add_batch(typename R::State &s, const typename R::In (&v)[B], uint32_t mask,
                                          i64 idx0, i64 kstride) {
    if constexpr (R::BATCHED) {
        R::template add_batch<B, V, FULL>(s, v, mask, idx0, kstride);
    } else {
#pragma unroll
        for (int b = 0; b < B; b++)
            if (FULL || ((mask >> b) & 1u)) R::add(s, v[b], idx0 + (i64)(b / V) * kstride + (b % V));
    }
}

template <typename R>
__device__ __forceinline__ void emit(const typename R::State &s, const RedOut &o, i64 j) {
    if (o.emit_state) {
        u64 w[3];
        R::pack(s, w);
#pragma unroll
        for (int k = 0; k < kStateWords; k++) o.states[(o.part * kStateWords + k) * o.outs + j] = w[k];
    } else {
        R::finalize(s, o, j);
    }
}

// 16-byte streaming load of V = 16 / sizeof(T) elements
template <typename T>
__device__ __forceinline__ void load16(const T *p, T *dst) {
    const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(p));
    *reinterpret_cast<uint4 *>(dst) = q;
}

// ---------------------------------------------------------------------- rows_cta (inner == 1)
// grid.x = rows * segs.  The CTA reduces positions [seg * seg_len, min(n, ...)) of one row.
template <typename R>
__global__ void __launch_bounds__(kRedThreads, R::MIN_CTAS) red_rows_cta_kernel(const typename R::In *__restrict__ a, RedOut o,
                                                                    i64 n, i64 segs, i64 seg_len, i64 index_offset) {
    using T = typename R::In;
    using State = typename R::State;
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int U = 4;
    constexpr int B = U * V;
    const int tid = threadIdx.x;
    const i64 row = blockIdx.x / segs, seg = blockIdx.x % segs;
    const i64 lo = seg * seg_len;
    const i64 hi = lo + seg_len < n ? lo + seg_len : n;
    const T *p = a + row * n;
    State s = R::init();
    if (hi > lo) {
        const int mis = (int)(((uintptr_t)(p + lo) & 15) / sizeof(T));
        i64 head = mis ? V - mis : 0;
        if (head > hi - lo) head = hi - lo;
        if (tid < head) R::add(s, p[lo + tid], index_offset + lo + tid);
        const i64 vlo = lo + head;
        const i64 nvec = (hi - vlo) / V;
        const T *pv = p + vlo;
        i64 j = tid;
        for (; j + (U - 1) * kRedThreads < nvec; j += U * kRedThreads) {
            alignas(16) T v[B];
#pragma unroll
            for (int u = 0; u < U; u++) load16(pv + (j + u * kRedThreads) * V, v + u * V);
            add_batch<R, B, V, true>(s, v, 0xffffffffu, index_offset + vlo + j * V, (i64)kRedThreads * V);
        }
        for (; j < nvec; j += kRedThreads) {
            alignas(16) T v[V];
            load16(pv + j * V, v);
            add_batch<R, V, V, true>(s, v, 0xffffffffu, index_offset + vlo + j * V, 0);
        }
        const i64 tlo = vlo + nvec * V;
        if (tid < hi - tlo) R::add(s, p[tlo + tid], index_offset + tlo + tid);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) R::merge(s, R::shfl(s, m));
    __shared__ State sm[kRedThreads / 32];
    if ((tid & 31) == 0) sm[tid >> 5] = s;
    __syncthreads();
    if (tid < 32) {
        s = tid < kRedThreads / 32 ? sm[tid] : R::init();
#pragma unroll
        for (int m = (kRedThreads / 32) / 2; m >= 1; m >>= 1) R::merge(s, R::shfl(s, m));
        if (tid == 0) {
            RedOut oo = o;
            if (segs > 1) oo.part = seg;
            emit<R>(s, oo, row);
        }
    }
}

// ------------------------------------------------------------------ group (G lanes / output)
// Output j = (o, c) reduces a[o, :, c]: element i sits at base + i * inner.  G > 1 only with
// inner == 1: the lanes of a group then read consecutive VEC-element vectors, so one group
// moves G * VEC * sizeof(T) contiguous bytes per load (VEC > 1 needs rows that start on a
// vector boundary: base aligned and n % VEC == 0).
template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T *p, T *dst) {
    if constexpr (VEC == 1) {
        dst[0] = *p;
    } else if constexpr (VEC * sizeof(T) == 8) {
        *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(p);
    } else {
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(p);
    }
}

template <typename R, int VEC>
__global__ void __launch_bounds__(kRedThreads, R::MIN_CTAS) red_group_kernel(const typename R::In *__restrict__ a, RedOut o,
                                                                    i64 n, i64 inner, int G, i64 index_offset) {
    using T = typename R::In;
    using State = typename R::State;
    constexpr int BV = VEC == 1 ? 8 : 4;  // vectors per batch
    constexpr int B = BV * VEC;
    const i64 gid = (i64)blockIdx.x * kRedThreads + threadIdx.x;
    const i64 j = gid / G;
    const int g = (int)(gid % G);
    const bool active = j < o.outs;
    const int nv = active ? (int)(n / VEC) : 0;  // n <= 4096 here: 32-bit positions
    const i64 jo = active ? j : 0;
    const T *p = a + (jo / inner) * n * inner + (jo % inner);
    const i64 estride = VEC == 1 ? inner : 1;  // element stride (VEC > 1 implies inner == 1)
    State s = R::init();
    int iv = g;
    for (; iv + (BV - 1) * G < nv; iv += BV * G) {
        alignas(16) T v[B];
#pragma unroll
        for (int u = 0; u < BV; u++) load_vec<T, VEC>(p + (i64)(iv + u * G) * VEC * estride, v + u * VEC);
        add_batch<R, B, VEC, true>(s, v, 0xffffffffu, index_offset + iv * VEC, (i64)G * VEC);
    }
    if (iv < nv) {
        // the last, partial batch: missing float elements become NaN (every reducer skips
        // NaN), so the mask-free path serves; integers carry a mask
        alignas(16) T v[B];
        uint32_t mask = 0;
#pragma unroll
        for (int u = 0; u < BV; u++) {
            const bool in = iv + u * G < nv;
            if (in) {
                load_vec<T, VEC>(p + (i64)(iv + u * G) * VEC * estride, v + u * VEC);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    if constexpr (std::is_integral<T>::value)
                        v[u * VEC + e] = T(0);
                    else
                        v[u * VEC + e] = quiet_nan<T>();
                }
            }
            mask |= (in ? ((1u << VEC) - 1u) : 0u) << (u * VEC);
        }
        if constexpr (std::is_integral<T>::value)
            add_batch<R, B, VEC, false>(s, v, mask, index_offset + iv * VEC, (i64)G * VEC);
        else
            add_batch<R, B, VEC, true>(s, v, 0xffffffffu, index_offset + iv * VEC, (i64)G * VEC);
    }
    for (int m = G >> 1; m >= 1; m >>= 1) R::merge(s, R::shfl(s, m));
    if (active && g == 0) emit<R>(s, o, j);
}

// ------------------------------------------------------------- rows_tile (inner == 1, short rows)
// Short rows cannot be read efficiently one row per thread or per sub-warp: a warp's load
// then touches many separate 32-byte pieces.  Here the CTA takes a CONTIGUOUS run of rows
// (<= 32 KB), brings it in with one bulk async copy (TMA engine, full-width DRAM bursts
// whatever n is), and G lanes per row reduce it out of shared memory.  With one lane per row
// and an even n the lanes start at rotated positions so that they fall on distinct banks.
constexpr int kTileBytes = 32 * 1024;

template <typename R>
__global__ void __launch_bounds__(kRedThreads, 6) red_rows_tile_kernel(const typename R::In *__restrict__ a,
                                                                                  RedOut o, int n, i64 rows, int tile_rows,
                                                                                  int log2g, i64 index_offset) {
    using T = typename R::In;
    using State = typename R::State;
    extern __shared__ __align__(128) unsigned char red_smem[];
    T *s = reinterpret_cast<T *>(red_smem);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const i64 row0 = (i64)blockIdx.x * tile_rows;
    const int nrows = (int)(rows - row0 < tile_rows ? rows - row0 : tile_rows);
    const int elems = nrows * n;
    const T *src = a + row0 * n;
    const uint32_t bytes = ((uint32_t)elems * (uint32_t)sizeof(T)) & ~15u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && bytes) {
        mbar_arrive_expect_tx(&bar, bytes);
        bulk_g2s(s, src, bytes, &bar);
    }
    const int tail0 = (int)(bytes / sizeof(T));
    if (tid < elems - tail0) s[tail0 + tid] = src[tail0 + tid];
    if (bytes) mbar_wait(&bar, 0);
    __syncthreads();

    const int G = 1 << log2g;
    const int rpp = kRedThreads >> log2g;  // rows per pass
    const int g = tid & (G - 1), rl = tid >> log2g;
    // one lane per row and an even n: lane l starts at element l % n, which spreads the
    // lanes over the banks (rpp is a multiple of 32, so the offset is fixed per thread)
    const int rot = (!R::ORDERED && G == 1 && (n & 1) == 0) ? (tid & 31) % n : 0;
    for (int r0 = 0; r0 < nrows; r0 += rpp) {
        const int r = r0 + rl;
        const bool act = r < nrows;
        const T *row = s + (act ? r : 0) * n;
        const int nn = act ? n : 0;
        State st = R::init();
        // f(value, element index) over this lane's elements g, g + G, ...
        auto lane_elements = [&](auto &&f) {
            if (rot) {  // G == 1: all n elements, starting at `rot`, wrapping once
                int pos = rot;
#pragma unroll 2
                for (int e = 0; e < nn; e++) {
                    f(row[pos], pos);
                    pos = pos + 1 == n ? 0 : pos + 1;
                }
            } else {
                const T *q = row + g;
#pragma unroll 4
                for (int e = g; e < nn; e += G, q += G) f(*q, e);
            }
        };
        if constexpr (R::TWO_PASS) {
            // the tile sits in shared memory, so the reference's own two loops (funcs.py:
            // 118-133) cost no second DRAM read: sum & count, then centred squares
            double sum = 0.0;
            int c = 0;
            lane_elements([&](T v, int) {
                if (v == v) {
                    sum += (double)v;
                    c += 1;
                }
            });
            for (int m = G >> 1; m >= 1; m >>= 1) {
                sum += shfl_xor_t(sum, m);
                c += shfl_xor_t(c, m);
            }
            const double mean = c > 0 ? sum * fast_rcp((double)c) : 0.0;
            double m2 = 0.0;
            lane_elements([&](T v, int) {
                const double d = (double)v - mean;
                if (v == v) m2 = fma(d, d, m2);
            });
            for (int m = G >> 1; m >= 1; m >>= 1) m2 += shfl_xor_t(m2, m);
            st.c = c;
            st.mean = mean;
            st.m2 = m2;
        } else if constexpr (R::ORDERED) {
            // nanarg*: 32-bit position inside the row while scanning, widened once
            using K = decltype(st.key);
            K key = st.key;
            int at = -1;
            lane_elements([&](T v, int e) {
                const K k = (K)v;
                const bool take = R::better(k, key) || (at < 0 && k == k);
                key = take ? k : key;
                at = take ? e : at;
            });
            st.key = key;
            st.idx = at < 0 ? -1 : index_offset + at;
            for (int m = G >> 1; m >= 1; m >>= 1) R::merge(st, R::shfl(st, m));
        } else {
            lane_elements([&](T v, int e) { R::add(st, v, index_offset + e); });
            for (int m = G >> 1; m >= 1; m >>= 1) R::merge(st, R::shfl(st, m));
        }
        if (act && g == 0) emit<R>(st, o, row0 + r);
    }
}

// ------------------------------------------------------------------------- cols (inner > 1)
// grid = (outer * segs, column tiles).  Thread tid < tprime = w * rps handles column
// col0 + tid % w and rows lo + tid / w + k * rps: consecutive threads, consecutive addresses.
template <typename R>
__global__ void __launch_bounds__(kRedThreads, R::MIN_CTAS) red_cols_kernel(const typename R::In *__restrict__ a, RedOut o,
                                                                   i64 n, i64 inner, i64 segs, i64 seg_len, int w,
                                                                   int rps, i64 index_offset) {
    using T = typename R::In;
    using State = typename R::State;
    constexpr int B = (sizeof(T) == 4 && R::MIN_CTAS >= 6 && !std::is_integral<T>::value) ? 16 : 8;
    const int tid = threadIdx.x;
    const i64 oi = blockIdx.x / segs, seg = blockIdx.x % segs;
    const i64 col = (i64)blockIdx.y * w + tid % w;
    const int r0 = tid / w;
    const bool active = tid < w * rps && col < inner;
    const i64 lo = seg * seg_len;
    i64 hi = lo + seg_len < n ? lo + seg_len : n;
    if (!active) hi = lo;
    const T *p = a + oi * n * inner + (active ? col : 0);
    State s = R::init();
    i64 i = lo + r0;
    for (; i + (i64)(B - 1) * rps < hi; i += (i64)B * rps) {
        T v[B];
#pragma unroll
        for (int b = 0; b < B; b++) v[b] = __ldcs(p + (i + (i64)b * rps) * inner);
        add_batch<R, B, 1, true>(s, v, 0xffffffffu, index_offset + i, rps);
    }
    if (i < hi) {
        T v[B];
        uint32_t mask = 0;
#pragma unroll
        for (int b = 0; b < B; b++) {
            const bool in = i + (i64)b * rps < hi;
            v[b] = in ? __ldcs(p + (i + (i64)b * rps) * inner) : T(0);
            mask |= (in ? 1u : 0u) << b;
        }
        add_batch<R, B, 1, false>(s, v, mask, index_offset + i, rps);
    }
    if (rps > 1) {
        __shared__ State sm[kRedThreads];
        sm[tid] = s;
        __syncthreads();
        if (r0 == 0)
            for (int r = 1; r < rps; r++) R::merge(s, sm[tid + r * w]);
    }
    if (active && r0 == 0) {
        RedOut oo = o;
        if (segs > 1) oo.part = seg;
        emit<R>(s, oo, oi * inner + col);
    }
}

// ----------------------------------------------------------------------------------- stream
// The main kernel for long reductions: rows (inner == 1, w = 1), narrow columns (inner <=
// 256, w = inner) and 16-byte-aligned wide columns (w = 256 of inner).  The CTA owns rows
// [lo, hi) x w columns of one outer slice and pulls them through a 3-stage shared-memory ring
// with bulk async copies (one copy per chunk when the tile is contiguous, one per row piece
// otherwise), so the bytes in flight live in shared memory, not in registers, and DRAM
// latency is hidden whatever the reducer costs.  Thread tid < w * rps reads ring positions
// tid + b * (w * rps): consecutive threads, consecutive words, always the same column.
// Threads that share a column merge through a shared-memory tree at the end.
constexpr int kStreamStages = 3;
constexpr int kStreamCtasPerSM = 4;

template <typename R>
__global__ void __launch_bounds__(kRedThreads, kStreamCtasPerSM) red_stream_kernel(
    const typename R::In *__restrict__ a, RedOut o, i64 n, i64 inner, i64 segs, i64 seg_len, int w, int rps,
    i64 index_offset) {
    using T = typename R::In;
    using State = typename R::State;
    constexpr int EPT = sizeof(T) == 4 ? 16 : 8;  // elements per thread per chunk (<= 16 KB chunks)
    extern __shared__ __align__(128) unsigned char red_smem[];
    __shared__ uint64_t full[kStreamStages];
    __shared__ State sm[kRedThreads];
    const int tid = threadIdx.x;
    const i64 oi = blockIdx.x / segs, seg = blockIdx.x % segs;
    const i64 col0 = (i64)blockIdx.y * w;
    const int tprime = w * rps;
    const int cw = (int)(inner - col0 < w ? inner - col0 : w);  // live columns of this tile
    const int col = tid % w, r0 = tid / w;
    const bool active = tid < tprime && col < cw;
    const i64 lo = seg * seg_len;
    const i64 hi = lo + seg_len < n ? lo + seg_len : n;
    const i64 nrows = hi > lo ? hi - lo : 0;
    const int chunk_rows = EPT * rps;
    const i64 nchunks = (nrows + chunk_rows - 1) / chunk_rows;
    const T *region = a + (oi * n + lo) * inner + col0;
    const bool contiguous = w == inner;
    constexpr uint32_t STAGE_BYTES = EPT * kRedThreads * sizeof(T);

    if (tid == 0) {
        for (int st = 0; st < kStreamStages; st++) mbar_init(&full[st], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // all threads call; chunk k goes to stage k % kStreamStages
    auto produce = [&](i64 k) {
        if (k >= nchunks) return;
        const int stg = (int)(k % kStreamStages);
        T *dst = reinterpret_cast<T *>(red_smem + (size_t)stg * STAGE_BYTES);
        const i64 rb = k * chunk_rows;
        const int rc = (int)(nrows - rb < chunk_rows ? nrows - rb : chunk_rows);
        if (contiguous) {
            if (tid == 0) {
                const uint32_t bytes = ((uint32_t)rc * (uint32_t)w * (uint32_t)sizeof(T)) & ~15u;
                mbar_arrive_expect_tx(&full[stg], bytes);
                if (bytes) bulk_g2s(dst, region + rb * inner, bytes, &full[stg]);
            }
        } else {
            const uint32_t row_bytes = (uint32_t)cw * (uint32_t)sizeof(T);
            if (tid == 0) mbar_arrive_expect_tx(&full[stg], row_bytes * (uint32_t)rc);
            if (tid < rc) bulk_g2s(dst + (size_t)tid * w, region + (rb + tid) * inner, row_bytes, &full[stg]);
        }
    };
    for (int k = 0; k < kStreamStages; k++) produce(k);

    State s = R::init();
    for (i64 k = 0; k < nchunks; k++) {
        const int stg = (int)(k % kStreamStages);
        const T *src = reinterpret_cast<const T *>(red_smem + (size_t)stg * STAGE_BYTES);
        mbar_wait(&full[stg], (uint32_t)((k / kStreamStages) & 1));
        const i64 rb = k * chunk_rows;
        const int rc = (int)(nrows - rb < chunk_rows ? nrows - rb : chunk_rows);
        if (active) {
            T v[EPT];
            if (rc == chunk_rows) {
#pragma unroll
                for (int b = 0; b < EPT; b++) v[b] = src[tid + b * tprime];
                add_batch<R, EPT, 1, true>(s, v, 0xffffffffu, index_offset + lo + rb + r0, rps);
            } else {
                // last chunk: rows past the end are NaN (floats) or masked (ints); a row tail
                // the 16-byte-granular copy left behind (contiguous tiles) comes from global
                const int copied = contiguous ? (int)((((uint32_t)rc * w * sizeof(T)) & ~15u) / sizeof(T)) : rc * w;
                uint32_t mask = 0;
#pragma unroll
                for (int b = 0; b < EPT; b++) {
                    const int pos = tid + b * tprime;
                    const bool in = r0 + b * rps < rc;
                    T x;
                    if constexpr (std::is_integral<T>::value)
                        x = T(0);
                    else
                        x = quiet_nan<T>();
                    if (in) x = pos < copied ? src[pos] : region[rb * inner + pos];
                    v[b] = x;
                    mask |= (in ? 1u : 0u) << b;
                }
                if constexpr (std::is_integral<T>::value)
                    add_batch<R, EPT, 1, false>(s, v, mask, index_offset + lo + rb + r0, rps);
                else
                    add_batch<R, EPT, 1, true>(s, v, 0xffffffffu, index_offset + lo + rb + r0, rps);
            }
        }
        __syncthreads();  // everyone is done with this stage: refill it
        produce(k + kStreamStages);
    }

    // merge the rps threads of each column (tree over r0), owner r0 == 0 emits
    if (rps > 1) {
        sm[tid] = s;
        __syncthreads();
        int span = 1;
        while (span < rps) span <<= 1;
        for (int h = span >> 1; h >= 1; h >>= 1) {
            if (tid < tprime && r0 < h && r0 + h < rps) {
                State mine = sm[tid];
                R::merge(mine, sm[tid + h * w]);
                sm[tid] = mine;
            }
            __syncthreads();
        }
        s = sm[tid];
    }
    if (active && r0 == 0) {
        RedOut oo = o;
        if (segs > 1) oo.part = seg;
        emit<R>(s, oo, oi * inner + col0 + col);
    }
}

// --------------------------------------------------------------------------------- merge
// Fold `parts` state records per output (segments of one launch, or shards of several
// devices: states[(part * 3 + word) * outs + j]).  lanes = 1: a thread per output (coalesced
// over j); lanes = 32: a warp per output (few outputs, many parts).
template <typename R>
__global__ void __launch_bounds__(kRedThreads) red_merge_kernel(const u64 *__restrict__ states, i64 parts, RedOut o,
                                                                 int lanes) {
    using State = typename R::State;
    const i64 gid = (i64)blockIdx.x * kRedThreads + threadIdx.x;
    const i64 j = gid / lanes;
    const int g = (int)(gid % lanes);
    const bool active = j < o.outs;
    State s = R::init();
    if (active)
        for (i64 p = g; p < parts; p += lanes) {
            u64 w[3];
#pragma unroll
            for (int k = 0; k < kStateWords; k++) w[k] = states[(p * kStateWords + k) * o.outs + j];
            R::merge(s, R::unpack(w));
        }
    for (int m = lanes >> 1; m >= 1; m >>= 1) R::merge(s, R::shfl(s, m));
    if (active && g == 0) emit<R>(s, o, j);
}

// Integer inputs hold no NaN: allnan / anynan / nancount are constants of the shape.
__global__ void red_const_kernel(RedOut o, int mode, i64 n) {
    const i64 j = (i64)blockIdx.x * kRedThreads + threadIdx.x;
    if (j >= o.outs) return;
    if (o.emit_state) {
        o.states[(o.part * kStateWords + 0) * o.outs + j] = (u64)n;
        o.states[(o.part * kStateWords + 1) * o.outs + j] = 0;
        o.states[(o.part * kStateWords + 2) * o.outs + j] = 0;
    } else if (mode == 0) {
        ((uint8_t *)o.out)[j] = n == 0;
    } else if (mode == 1) {
        ((uint8_t *)o.out)[j] = 0;
    } else {
        ((i64 *)o.out)[j] = n;
    }
}

// ------------------------------------------------------------------------------ geometry
enum RedMode { RED_ROWS_CTA = 0, RED_GROUP = 1, RED_COLS = 2, RED_ROWS_TILE = 3, RED_STREAM = 4 };
struct RedGeom {
    int mode;
    i64 segs, seg_len;
    int G, vec;    // group
    int w, rps;    // cols
    i64 coltiles;  // cols
    int tile_rows;  // rows_tile
};

// CTAs of one full wave for an op (its reducer's MIN_CTAS): segment counts aim at exactly that
inline i64 target_ctas(int op) {
    const int per_sm = (op == NBG_RED_NANVAR || op == NBG_RED_NANSTD) ? 4 : (op == NBG_RED_NANARGMAX || op == NBG_RED_NANARGMIN) ? 5 : 6;
    return (i64)kNumSMs * per_sm;
}

inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// itemsize / addr: element size and address of the input (vector width of the group kernel)
constexpr i64 kStreamWave = (i64)kNumSMs * kStreamCtasPerSM;

// Segments per slice for `slices` independent slices: with few slices, as many as fill ONE
// wave of resident CTAs (equal work, no tail); otherwise at least four waves in total.
inline i64 pick_segs(i64 slices, i64 wave, i64 max_segs) {
    if (slices < 1) slices = 1;
    if (max_segs < 1) max_segs = 1;
    if (slices * 4 <= wave) {
        const i64 segs = wave / slices;
        return segs > max_segs ? max_segs : segs;
    }
    if (slices >= 8 * wave) return 1;
    // a few candidates past four waves: take the one whose last wave is fullest
    i64 best = 1;
    double best_eff = -1.0;
    const i64 s0 = ceil_div(4 * wave, slices);
    for (i64 segs = s0; segs < s0 + 4; segs++) {
        const i64 sg = segs > max_segs ? max_segs : segs;
        const i64 ctas = slices * sg;
        // fuller last wave, but every extra segment costs a prologue and a merge record
        const double eff = (double)ctas / (double)(ceil_div(ctas, wave) * wave) - 0.02 * (double)(segs - s0);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = sg;
        }
    }
    return best;
}

RedGeom red_geometry(int op, i64 outer, i64 n, i64 inner, int itemsize = 4, uintptr_t addr = 0) {
    const i64 kTargetCtas = target_ctas(op);
    RedGeom g{};
    g.segs = 1;
    g.seg_len = n > 0 ? n : 1;
    g.G = 1;
    g.vec = 1;
    if (inner == 1) {
        if (n > 0 && n * itemsize <= 4096 && outer >= 64 && addr % 16 == 0) {
            g.mode = RED_ROWS_TILE;
            g.tile_rows = (int)((kTileBytes / (n * itemsize)) & ~(i64)3);
            int G = 1;
            while (G < 32 && (i64)G * 16 < n) G <<= 1;
            g.G = G;
            return g;
        }
        if (n <= 4096) {
            g.mode = RED_GROUP;
            int vec = 16 / itemsize;
            while (vec > 1 && (n % vec != 0 || addr % (uintptr_t)(vec * itemsize) != 0)) vec >>= 1;
            g.vec = vec;
            const i64 nv = n / vec;  // each lane should own >= 2 vectors
            int G = 1;
            while (G < 32 && (i64)G * 4 <= nv) G <<= 1;
            g.G = G;
            return g;
        }
        // (the stream kernel handles rows too -- w = 1 -- but measured ~8% slower than direct
        // 16-byte loads on this fully coalesced shape; NBG_RED_STREAM_ROWS=1 selects it)
        static const bool stream_rows = getenv("NBG_RED_STREAM_ROWS") != nullptr;
        const bool stream = stream_rows && addr % 16 == 0 && (outer == 1 || (n * itemsize) % 16 == 0);
        g.mode = stream ? RED_STREAM : RED_ROWS_CTA;
        g.w = 1;
        g.rps = kRedThreads;
        g.coltiles = 1;
        const i64 segs = pick_segs(outer, stream ? kStreamWave : kTargetCtas, ceil_div(n, 16384));
        i64 seg_len = ceil_div(ceil_div(n, segs), 1024) * 1024;
        g.seg_len = seg_len;
        g.segs = ceil_div(n, seg_len);
        return g;
    }
    const i64 w = inner <= kRedThreads ? inner : kRedThreads;
    if (n * w < 2048) {  // too little work per (outer, column tile) for a CTA: thread per output
        g.mode = RED_GROUP;
        return g;
    }
    g.mode = RED_COLS;
    g.w = (int)w;
    g.rps = inner <= kRedThreads ? (int)(kRedThreads / inner) : 1;
    g.coltiles = ceil_div(inner, w);
    const i64 base = g.coltiles * (outer > 0 ? outer : 1);
    // bulk copies need 16-byte aligned tiles: every outer slice and every row piece
    const bool stream = addr % 16 == 0 && (inner * itemsize) % 16 == 0 ||
                        (addr % 16 == 0 && inner <= kRedThreads && (n * inner * itemsize) % 16 == 0);
    if (stream) {
        // worth it only if a segment spans enough ring chunks to amortise the pipeline fill
        const i64 chunk_rows = (i64)(itemsize == 4 ? 16 : 8) * g.rps;
        const i64 segs = pick_segs(base, kStreamWave, n / (8 * chunk_rows));
        const i64 seg_len = ceil_div(ceil_div(n, segs), 4) * 4;  // segment starts stay 16-byte aligned
        if (seg_len >= 8 * chunk_rows || seg_len >= n) {
            if (n >= 6 * chunk_rows) {
                g.mode = RED_STREAM;
                g.seg_len = seg_len;
                g.segs = ceil_div(n, seg_len);
                return g;
            }
        }
    }
    const i64 segs = pick_segs(base, kTargetCtas, ceil_div(n, (i64)16 * g.rps));
    g.seg_len = ceil_div(n, segs);
    g.segs = ceil_div(n, g.seg_len);
    return g;
}

// ------------------------------------------------------------------------------ launching
struct RedArgs {
    int op;
    const void *a;
    void *out;
    u64 *states;  // partial mode: caller's record array (part 0 of 1)
    i64 outer, n, inner, ddof, index_offset;
    void *workspace;
    size_t workspace_bytes;
    cudaStream_t stream;
    bool partial;
};

template <typename R>
int launch_merge(const u64 *states, i64 parts, const RedOut &o, cudaStream_t stream) {
    if (o.outs <= 0) return NBG_OK;
    const int lanes = (o.outs < 2048 && parts > 4) ? 32 : 1;
    const i64 threads = o.outs * lanes;
    red_merge_kernel<R><<<(unsigned)ceil_div(threads, kRedThreads), kRedThreads, 0, stream>>>(states, parts, o, lanes);
    return check_launch("nbg_reduce merge");
}

template <typename R>
int launch_reduce(const RedArgs &x) {
    using T = typename R::In;
    const i64 outs = x.outer * x.inner;
    if (outs <= 0) return NBG_OK;
    const RedGeom g = red_geometry(x.op, x.outer, x.n, x.inner, (int)sizeof(T), (uintptr_t)x.a);
    RedOut fin{x.out, x.states, outs, 0, x.partial ? 1 : 0, x.n, x.ddof};
    RedOut first = fin;
    if (g.segs > 1) {
        const size_t need = (size_t)g.segs * kStateWords * outs * sizeof(u64);
        if (x.workspace == nullptr || x.workspace_bytes < need)
            return fail(NBG_ERR_WORKSPACE, "nbg_reduce: workspace smaller than nbg_reduce_workspace_bytes()");
        first.states = (u64 *)x.workspace;
        first.emit_state = 1;
    }
    const T *a = (const T *)x.a;
    int rc;
    if (g.mode == RED_STREAM) {
        const i64 gx = x.outer * g.segs;
        if (gx > 0x7fffffff || g.coltiles > 65535) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: shape too large");
        if (int rc2 = allow_big_smem(red_stream_kernel<R>, "nbg_reduce stream smem")) return rc2;
        constexpr size_t kSmem = (size_t)kStreamStages * (sizeof(T) == 4 ? 16 : 8) * kRedThreads * sizeof(T);
        red_stream_kernel<R><<<dim3((unsigned)gx, (unsigned)g.coltiles), kRedThreads, kSmem, x.stream>>>(
            a, first, x.n, x.inner, g.segs, g.seg_len, g.w, g.rps, x.index_offset);
        rc = check_launch("nbg_reduce stream");
    } else if (g.mode == RED_ROWS_CTA) {
        const i64 ctas = x.outer * g.segs;
        if (ctas > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: too many rows");
        red_rows_cta_kernel<R><<<(unsigned)ctas, kRedThreads, 0, x.stream>>>(a, first, x.n, g.segs, g.seg_len,
                                                                             x.index_offset);
        rc = check_launch("nbg_reduce rows");
    } else if (g.mode == RED_ROWS_TILE) {
        const i64 ctas = ceil_div(x.outer, g.tile_rows);
        if (ctas > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: too many rows");
        red_rows_tile_kernel<R><<<(unsigned)ctas, kRedThreads, (size_t)g.tile_rows * x.n * sizeof(T), x.stream>>>(
            a, first, (int)x.n, x.outer, g.tile_rows, __builtin_ctz((unsigned)g.G), x.index_offset);
        rc = check_launch("nbg_reduce rows_tile");
    } else if (g.mode == RED_GROUP) {
        const i64 ctas = ceil_div(outs * g.G, kRedThreads);
        if (ctas > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: too many outputs");
        constexpr int VMAX = 16 / (int)sizeof(T);
        if (g.vec == VMAX)
            red_group_kernel<R, VMAX><<<(unsigned)ctas, kRedThreads, 0, x.stream>>>(a, first, x.n, x.inner, g.G,
                                                                                  x.index_offset);
        else if (g.vec == 2 && VMAX > 2)
            red_group_kernel<R, (VMAX > 2 ? 2 : 1)><<<(unsigned)ctas, kRedThreads, 0, x.stream>>>(
                a, first, x.n, x.inner, g.G, x.index_offset);
        else
            red_group_kernel<R, 1><<<(unsigned)ctas, kRedThreads, 0, x.stream>>>(a, first, x.n, x.inner, g.G,
                                                                               x.index_offset);
        rc = check_launch("nbg_reduce group");
    } else {
        const i64 gx = x.outer * g.segs;
        if (gx > 0x7fffffff || g.coltiles > 65535) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: shape too large");
        red_cols_kernel<R><<<dim3((unsigned)gx, (unsigned)g.coltiles), kRedThreads, 0, x.stream>>>(
            a, first, x.n, x.inner, g.segs, g.seg_len, g.w, g.rps, x.index_offset);
        rc = check_launch("nbg_reduce cols");
    }
    if (rc) return rc;
    if (g.segs > 1) return launch_merge<R>((const u64 *)x.workspace, g.segs, fin, x.stream);
    return NBG_OK;
}

// op x dtype -> reducer.  F is a generic callable taking a reducer tag.
template <typename R>
struct Tag {
    using type = R;
};

template <typename T, typename F>
int with_reducer_t(int op, F &&f) {
    constexpr bool is_int = std::is_integral<T>::value;
    switch (op) {
        case NBG_RED_ALLNAN: return f(Tag<RCount<T, 0>>{});
        case NBG_RED_ANYNAN: return f(Tag<RCount<T, 1>>{});
        case NBG_RED_NANCOUNT: return f(Tag<RCount<T, 2>>{});
        case NBG_RED_NANSUM: return f(Tag<RSum<T>>{});
        case NBG_RED_NANARGMAX: return f(Tag<RArg<T, true>>{});
        case NBG_RED_NANARGMIN: return f(Tag<RArg<T, false>>{});
        case NBG_RED_NANMAX: return f(Tag<RExt<T, true>>{});
        case NBG_RED_NANMIN: return f(Tag<RExt<T, false>>{});
        default: break;
    }
    if constexpr (!is_int) {
        switch (op) {
            case NBG_RED_NANMEAN: return f(Tag<RMean<T>>{});
            case NBG_RED_NANVAR: return f(Tag<RVar<T, false>>{});
            case NBG_RED_NANSTD: return f(Tag<RVar<T, true>>{});
            default: break;
        }
    } else if (op == NBG_RED_NANMEAN || op == NBG_RED_NANVAR || op == NBG_RED_NANSTD) {
        return fail(NBG_ERR_BAD_DTYPE, "nbg_reduce: nanmean/nanvar/nanstd take float32/float64 (cast integers first)");
    }
    return fail(NBG_ERR_BAD_OP, "nbg_reduce: unknown op");
}

template <typename F>
int with_reducer(int op, int dtype, F &&f) {
    switch (dtype) {
        case NBG_F32: return with_reducer_t<float>(op, f);
        case NBG_F64: return with_reducer_t<double>(op, f);
        case NBG_I32: return with_reducer_t<int32_t>(op, f);
        case NBG_I64: return with_reducer_t<int64_t>(op, f);
        default: return fail(NBG_ERR_BAD_DTYPE, "nbg_reduce: dtype must be NBG_F32|F64|I32|I64");
    }
}

int check_shape(i64 outer, i64 n, i64 inner) {
    if (outer < 0 || n < 0 || inner < 0) return fail(NBG_ERR_BAD_ARG, "nbg_reduce: negative extent");
    return NBG_OK;
}

int reduce_entry(int op, int dtype, const RedArgs &x) {
    if (int rc = check_shape(x.outer, x.n, x.inner)) return rc;
    const bool is_int = dtype == NBG_I32 || dtype == NBG_I64;
    if (is_int && op >= NBG_RED_ALLNAN && op <= NBG_RED_NANCOUNT) {
        const i64 outs = x.outer * x.inner;
        if (outs <= 0) return NBG_OK;
        RedOut o{x.out, x.states, outs, 0, x.partial ? 1 : 0, x.n, x.ddof};
        red_const_kernel<<<(unsigned)ceil_div(outs, kRedThreads), kRedThreads, 0, x.stream>>>(o, op, x.n);
        return check_launch("nbg_reduce const");
    }
    return with_reducer(op, dtype, [&](auto tag) {
        using R = typename decltype(tag)::type;
        return launch_reduce<R>(x);
    });
}

}  // namespace
}  // namespace nbg

using namespace nbg;

extern "C" size_t nbg_reduce_workspace_bytes(int op, int dtype, int64_t outer, int64_t n, int64_t inner) {
    if (outer <= 0 || n <= 0 || inner <= 0) return 0;
    const int itemsize = (dtype == NBG_F64 || dtype == NBG_I64) ? 8 : 4;
    // the kernel choice depends on the input's alignment, unknown here: size for either
    const i64 s0 = red_geometry(op, outer, n, inner, itemsize, 0).segs;
    const i64 s1 = red_geometry(op, outer, n, inner, itemsize, 4).segs;
    const i64 segs = s0 > s1 ? s0 : s1;
    if (segs <= 1) return 0;
    return (size_t)segs * kStateWords * (size_t)(outer * inner) * sizeof(u64);
}

extern "C" int nbg_reduce(int op, int dtype, const void *a, void *out, int64_t outer, int64_t n, int64_t inner,
                          int64_t ddof, void *workspace, size_t workspace_bytes, void *stream) {
    RedArgs x{op, a, out, nullptr, outer, n, inner, ddof, 0, workspace, workspace_bytes, (cudaStream_t)stream, false};
    if (outer * inner > 0 && (out == nullptr || (a == nullptr && n > 0)))
        return fail(NBG_ERR_BAD_ARG, "nbg_reduce: null pointer");
    return reduce_entry(op, dtype, x);
}

extern "C" int nbg_reduce_partial(int op, int dtype, const void *a, void *states, int64_t outer, int64_t n,
                                  int64_t inner, int64_t index_offset, void *workspace, size_t workspace_bytes,
                                  void *stream) {
    RedArgs x{op, a,         nullptr,         (u64 *)states,        outer, n,   inner, 0, index_offset,
              workspace, workspace_bytes, (cudaStream_t)stream, true};
    if (outer * inner > 0 && (states == nullptr || (a == nullptr && n > 0)))
        return fail(NBG_ERR_BAD_ARG, "nbg_reduce_partial: null pointer");
    return reduce_entry(op, dtype, x);
}

extern "C" int nbg_reduce_merge(int op, int dtype, const void *states, int64_t parts, int64_t outs, void *out,
                                int64_t n_total, int64_t ddof, void *stream) {
    if (parts < 0 || outs < 0) return fail(NBG_ERR_BAD_ARG, "nbg_reduce_merge: negative extent");
    if (outs == 0) return NBG_OK;
    if (out == nullptr || (states == nullptr && parts > 0))
        return fail(NBG_ERR_BAD_ARG, "nbg_reduce_merge: null pointer");
    RedOut o{out, nullptr, outs, 0, 0, n_total, ddof};
    const bool is_int = dtype == NBG_I32 || dtype == NBG_I64;
    // counts of integer shards are plain counts too: fold them with the float counter
    if (is_int && op >= NBG_RED_ALLNAN && op <= NBG_RED_NANCOUNT) dtype = NBG_F64;
    return with_reducer(op, dtype, [&](auto tag) {
        using R = typename decltype(tag)::type;
        return launch_merge<R>((const u64 *)states, parts, o, (cudaStream_t)stream);
    });
}
