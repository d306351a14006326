// nbg_group.cu -- grouped NaN-aware reductions (group_nan*) for sm_100a.
//
// Replaces the scatter-reduce loops of numbagg/grouped.py:7-270 (dispatched by
// groupndreduce, numbagg/decorators.py:490-674).
//
// Workspace ("group workspace"): one record of 1, 2 or 4 eight-byte slots per (row, label),
// ws[row][label][slot] (ws_layout() below).  Multi-channel ops keep their channels in ONE
// 16/32-byte record, so the 2-3 atomics an element issues for mean/var/std/arg*/first/last land
// in the same sector (2-3x less L2/DRAM traffic than a channel-major layout when labels are
// random and the table exceeds L2); single-channel ops keep an 8-byte-per-label table.
// Channels (the slot a channel occupies depends on the op):
//   op                         ch0                      ch1                 ch2
//   nansum / sum_of_squares    sum   (f64 | i64)        -                   -
//   nanmean                    sum                      -                   count (i64)
//   nancount                   -                        -                   count
//   nanvar / nanstd            sum                      sum of squares      count
//   nanprod                    product (f64 | i64)      -                   -
//   nanmin / nanmax            ordered key (u64, 0=empty; min uses the inverted key)
//   nanany / nanall            flag (i64)
//   nanfirst / nanlast         raw value bits           global flat index (i64)
//   nanargmax / nanargmin      ordered key              global flat index of the FIRST extreme
// Float data accumulate in double (global RED.ADD.F64), integers in int64 (wrap-around, like
// the reference's in-dtype accumulation modulo 2^bits).  Comparisons for min/max/arg* happen
// on the value converted to double, exactly like the reference (its scratch arrays are
// float64: grouped.py:56, 80, 213, 231).
//
// Kernels
//   group_atomic_kernel      any shape / cardinality: one global atomic per element and
//                            channel.  HBM/L2-atomic bound; the high-cardinality path.
//   group_rowbins_kernel     labels shared by all rows, bins of a few rows fit in shared
//                            memory: one CTA owns (row group, column segment), streams row
//                            tiles through TMA bulk copies and accumulates into PRIVATE
//                            shared-memory bins that only one lane ever touches -- no
//                            atomics, and each bin sees its elements in ascending column
//                            order, i.e. the reference's own summation order.
#include "nbg_common.cuh"
#include "nbg_group_rowbins.cuh"
#include "nbg_group_rowbins2.cuh"
#include "nbg_group_partition.cuh"

namespace nbg {

constexpr int64_t kIdxNone = INT64_MAX;

template <typename V>
struct VTraits;
template <>
struct VTraits<float> {
    using Acc = double;
    static constexpr bool is_float = true;
    __device__ static __forceinline__ float nan_out() { return quiet_nan<float>(); }
};
template <>
struct VTraits<double> {
    using Acc = double;
    static constexpr bool is_float = true;
    __device__ static __forceinline__ double nan_out() { return quiet_nan<double>(); }
};
template <>
struct VTraits<int32_t> {
    using Acc = long long;
    static constexpr bool is_float = false;
    __device__ static __forceinline__ int32_t nan_out() { return INT32_MIN; }  // x86 cvttsd2si(NaN)
};
template <>
struct VTraits<int64_t> {
    using Acc = long long;
    static constexpr bool is_float = false;
    __device__ static __forceinline__ int64_t nan_out() { return INT64_MIN; }
};

// square / product in the VALUE type first (numba: V*V is V), then widened to the accumulator
__device__ __forceinline__ double sq_as_input(float v) { return (double)__fmul_rn(v, v); }
__device__ __forceinline__ double sq_as_input(double v) { return __dmul_rn(v, v); }
__device__ __forceinline__ long long sq_as_input(int32_t v) { return (long long)v * (long long)v; }
__device__ __forceinline__ long long sq_as_input(int64_t v) {
    return (long long)((unsigned long long)v * (unsigned long long)v);
}

// ---- L2 cache-policy hints -------------------------------------------------------------------
// The accumulator table is the only data of the per-element-label path that is touched more than
// once; its reductions carry an L2::evict_last policy (createpolicy) while the inputs stream
// through with evict-first loads (__ldcs), so the table survives in L2 next to a 32 GB stream.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
struct L2Hint {
    uint64_t pol;
    bool on;
};
__device__ __forceinline__ void red_add(double *p, double v, const L2Hint &h) {
    if (h.on) asm volatile("red.global.add.f64.L2::cache_hint [%0], %1, %2;" ::"l"(p), "d"(v), "l"(h.pol) : "memory");
    else atomicAdd(p, v);
}
__device__ __forceinline__ void red_add(long long *p, long long v, const L2Hint &h) {
    if (h.on) asm volatile("red.global.add.u64.L2::cache_hint [%0], %1, %2;" ::"l"(p), "l"(v), "l"(h.pol) : "memory");
    else atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}
__device__ __forceinline__ void red_max(unsigned long long *p, unsigned long long v, const L2Hint &h) {
    if (h.on) asm volatile("red.global.max.u64.L2::cache_hint [%0], %1, %2;" ::"l"(p), "l"(v), "l"(h.pol) : "memory");
    else atomicMax(p, v);
}
__device__ __forceinline__ void red_max(long long *p, long long v, const L2Hint &h) {
    if (h.on) asm volatile("red.global.max.s64.L2::cache_hint [%0], %1, %2;" ::"l"(p), "l"(v), "l"(h.pol) : "memory");
    else atomicMax(p, v);
}
__device__ __forceinline__ void red_min(long long *p, long long v, const L2Hint &h) {
    if (h.on) asm volatile("red.global.min.s64.L2::cache_hint [%0], %1, %2;" ::"l"(p), "l"(v), "l"(h.pol) : "memory");
    else atomicMin(p, v);
}
__device__ __forceinline__ unsigned long long ld_hint(const unsigned long long *p, const L2Hint &h) {
    if (!h.on) return *p;
    unsigned long long v;
    asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(h.pol) : "memory");
    return v;
}

__device__ __forceinline__ void atomic_add_acc(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_acc(long long *p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}
// Products accumulate with a CAS loop.  float32 products round (and overflow / underflow) in
// float32 after every factor, like the reference's in-dtype accumulation (grouped.py:124-132);
// the slot still holds a double.  Integers multiply modulo 2^64 (truncated at the end).
template <typename V>
__device__ __forceinline__ void atomic_mul_acc(void *p, V v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        unsigned long long next;
        if constexpr (std::is_same<V, float>::value) {
            const float prod = __fmul_rn((float)__longlong_as_double((long long)assumed), v);
            next = (unsigned long long)__double_as_longlong((double)prod);
        } else if constexpr (std::is_same<V, double>::value) {
            next = (unsigned long long)__double_as_longlong(__dmul_rn(__longlong_as_double((long long)assumed), v));
        } else {
            next = assumed * (unsigned long long)(long long)v;
        }
        old = atomicCAS(a, assumed, next);
    } while (old != assumed);
}

// Record layout per op: only the slots an op uses are materialised, so single-channel ops keep
// an 8-byte-per-label table (L2-resident up to ~15 M labels) while multi-channel ops keep their
// channels in one sector.  stride = 8-byte words per record; slot[c] = position of channel c.
struct WsLayout {
    int stride;   // 8-byte words between consecutive records
    int slot[3];  // word offset of channel c inside a record (interleaved) or plane index (planar)
    int planar;   // 1: channels are separate (rows*K)-word planes
    int words;    // total 8-byte words per (row, label)
};
// Tables larger than this cannot stay in L2 next to the input stream: multi-channel additive ops
// then switch to one plane per channel and the per-element-label path makes one pass per plane
// with that plane pinned in L2 (measured on config 5, 10 M labels: nanvar 97.8 ms interleaved ->
// see DESIGN.md 4.3).
constexpr size_t kL2TableBytes = (size_t)96 << 20;
__host__ __device__ inline WsLayout ws_layout(int op, int64_t slots = 0) {
    if ((op == NBG_GROUP_NANMEAN || op == NBG_GROUP_NANVAR || op == NBG_GROUP_NANSTD) &&
        (size_t)slots * (op == NBG_GROUP_NANMEAN ? 16 : 32) > kL2TableBytes) {
        return WsLayout{1, {0, 1, op == NBG_GROUP_NANMEAN ? 1 : 2}, 1, op == NBG_GROUP_NANMEAN ? 2 : 4};
    }
    switch (op) {
        case NBG_GROUP_NANCOUNT:
            return WsLayout{1, {0, 0, 0}, 0, 1};  // count only (ch2)
        case NBG_GROUP_NANMEAN:
            return WsLayout{2, {0, 0, 1}, 0, 2};  // sum, count: touched together -> one record
        case NBG_GROUP_NANVAR:
        case NBG_GROUP_NANSTD:
            return WsLayout{4, {0, 1, 2}, 0, 4};  // sum, sum of squares, count, (spare)
        case NBG_GROUP_NANFIRST:
        case NBG_GROUP_NANLAST:
        case NBG_GROUP_NANARGMAX:
        case NBG_GROUP_NANARGMIN:
            // key/bits plane + index plane: each pass touches ONE plane, which stays
            // L2-resident up to ~15 M labels (measured: 2x faster than 16-byte records)
            return WsLayout{1, {0, 1, 0}, 1, 2};
        default:
            return WsLayout{1, {0, 0, 0}, 0, 1};  // one accumulator / key / flag
    }
}
struct GroupWs {
    void *ch[3];     // ch[c] + record * stride (in 8-byte words)
    int64_t stride;
    __host__ __device__ static GroupWs carve(void *base, int op, int64_t rows, int64_t K) {
        GroupWs w;
        const WsLayout l = ws_layout(op, rows * K);
        const size_t plane = l.planar ? (size_t)rows * (size_t)K : 1;
        for (int c = 0; c < 3; c++) w.ch[c] = static_cast<unsigned char *>(base) + (size_t)l.slot[c] * plane * 8;
        w.stride = l.stride;
        return w;
    }
};

// ------------------------------------------------------------------------------------- init
__global__ void group_init_kernel(GroupWs ws, int op, int is_float, int64_t slots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const WsLayout l = ws_layout(op, slots);
    unsigned long long *w0 = reinterpret_cast<unsigned long long *>(ws.ch[0]) + i * ws.stride;
    unsigned long long *w1 = reinterpret_cast<unsigned long long *>(ws.ch[1]) + i * ws.stride;
    if (l.planar) {
        *w0 = 0;
        *w1 = 0;
        reinterpret_cast<unsigned long long *>(ws.ch[2])[i * ws.stride] = 0;  // third plane (var) or an alias
    } else {
        unsigned long long *rec = w0 - l.slot[0];
        for (int w = 0; w < l.stride; w++) rec[w] = 0;
    }
    switch (op) {
        case NBG_GROUP_NANPROD:
            *w0 = is_float ? (unsigned long long)__double_as_longlong(1.0) : 1ull;
            break;
        case NBG_GROUP_NANALL:
            *w0 = 1ull;
            break;
        case NBG_GROUP_NANFIRST:
        case NBG_GROUP_NANARGMAX:
        case NBG_GROUP_NANARGMIN:
            *w1 = (unsigned long long)kIdxNone;
            break;
        case NBG_GROUP_NANLAST:
            *w1 = (unsigned long long)(long long)-1;
            break;
        default:
            break;
    }
}

// --------------------------------------------------------------------- generic atomic kernel
// PHASE 1 of arg* re-reads the data and records the smallest index whose key equals the
// group's extreme; every other op has PHASE 0 only.
// CHM: bit c set = this pass updates channel c (planar multi-pass of mean / var / std; 7 = all).
template <typename V, typename L, int OP, int PHASE, int CHM = 7>
__global__ void __launch_bounds__(256) group_atomic_kernel(const V *__restrict__ values, const L *__restrict__ labels,
                                                           int labels_per_row, GroupWs ws, int64_t rows, int64_t n,
                                                           int64_t K, int64_t index_offset, int64_t blocks_per_row,
                                                           int l2_hint) {
    using Acc = typename VTraits<V>::Acc;
    L2Hint hint{0, l2_hint != 0};
    if (hint.on) hint.pol = l2_policy_evict_last();
    const int64_t row = blockIdx.x / blocks_per_row;
    const int64_t blk = blockIdx.x % blocks_per_row;
    const V *vrow = values + row * n;
    const L *lrow = labels + (labels_per_row ? row * n : 0);
    Acc *c0 = reinterpret_cast<Acc *>(ws.ch[0]) + row * K * ws.stride;
    Acc *c1 = reinterpret_cast<Acc *>(ws.ch[1]) + row * K * ws.stride;
    long long *cnt = reinterpret_cast<long long *>(ws.ch[2]) + row * K * ws.stride;
    unsigned long long *key0 = reinterpret_cast<unsigned long long *>(ws.ch[0]) + row * K * ws.stride;
    long long *idx1 = reinterpret_cast<long long *>(ws.ch[1]) + row * K * ws.stride;
    constexpr int PER = 8;
    const int64_t base = blk * (256 * PER);
    // All 16 loads of a thread are issued before the first one is used: with a label test between
    // the label load and the value load the kernel exposed two DRAM latencies per element and sat
    // in long_scoreboard at 43 % of the DRAM bandwidth (ncu, config 5).  Streaming loads
    // (evict-first): the inputs are read once and must not push the accumulator table out of L2,
    // where the atomics are resolved.
    int64_t labs[PER];
    V vals[PER];
    const bool whole = base + 256 * PER <= n;
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int64_t i = base + (int64_t)q * 256 + threadIdx.x;
        const bool in = whole || i < n;
        labs[q] = in ? (int64_t)__ldcs(lrow + i) : (int64_t)-1;
        vals[q] = in ? __ldcs(vrow + i) : (V)0;
    }
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int64_t i = base + (int64_t)q * 256 + threadIdx.x;
        int64_t label = labs[q];
        if (label < 0 || label >= K) continue;
        label *= ws.stride;  // record offset in 8-byte words
        const V v = vals[q];
        if (is_nan(v)) continue;
        const int64_t gi = index_offset + i;
        if (OP == NBG_GROUP_NANSUM) {
            red_add(c0 + label, (Acc)v, hint);
        } else if (OP == NBG_GROUP_NANMEAN) {
            if (CHM & 1) red_add(c0 + label, (Acc)v, hint);
            if (CHM & 4) red_add(cnt + label, 1ll, hint);
        } else if (OP == NBG_GROUP_NANCOUNT) {
            red_add(cnt + label, 1ll, hint);
        } else if (OP == NBG_GROUP_NANSUM_OF_SQUARES) {
            red_add(c0 + label, (Acc)sq_as_input(v), hint);
        } else if (OP == NBG_GROUP_NANVAR || OP == NBG_GROUP_NANSTD) {
            if (CHM & 1) red_add(c0 + label, (Acc)v, hint);
            if (CHM & 2) red_add(c1 + label, (Acc)sq_as_input(v), hint);
            if (CHM & 4) red_add(cnt + label, 1ll, hint);
        } else if (OP == NBG_GROUP_NANPROD) {
            atomic_mul_acc<V>(c0 + label, v);
        } else if (OP == NBG_GROUP_NANMAX) {
            red_max(key0 + label, order_key((double)v), hint);
        } else if (OP == NBG_GROUP_NANMIN) {
            red_max(key0 + label, ~order_key((double)v), hint);
        } else if (OP == NBG_GROUP_NANANY) {
            if (v != (V)0) reinterpret_cast<volatile long long *>(c0)[label] = 1;
        } else if (OP == NBG_GROUP_NANALL) {
            if (v == (V)0) reinterpret_cast<volatile long long *>(c0)[label] = 0;
        } else if (OP == NBG_GROUP_NANFIRST) {
            red_min(idx1 + label, (long long)gi, hint);
        } else if (OP == NBG_GROUP_NANLAST) {
            red_max(idx1 + label, (long long)gi, hint);
        } else if (OP == NBG_GROUP_NANARGMAX || OP == NBG_GROUP_NANARGMIN) {
            const unsigned long long k = OP == NBG_GROUP_NANARGMAX ? order_key((double)v) : ~order_key((double)v);
            if (PHASE == 0) {
                red_max(key0 + label, k, hint);
            } else {
                if (k == ld_hint(key0 + label, hint)) atomicMin(idx1 + label, (long long)gi);
            }
        }
    }
}

// first/last: fetch the winning value of this shard into ch0 (one thread per (row, label)).
template <typename V>
__global__ void group_gather_kernel(const V *__restrict__ values, GroupWs ws, int64_t rows, int64_t n, int64_t K,
                                    int64_t index_offset) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rows * K) return;
    const int64_t row = s / K;
    const long long gi = reinterpret_cast<long long *>(ws.ch[1])[s * ws.stride];
    const int64_t local = gi - index_offset;
    if (local < 0 || local >= n) return;  // winner lives in another shard (or none)
    const V v = values[row * n + local];
    unsigned long long bits;
    if (sizeof(V) == 8) {
        bits = *reinterpret_cast<const unsigned long long *>(&v);
    } else {
        bits = (unsigned long long)*reinterpret_cast<const unsigned int *>(&v);
    }
    reinterpret_cast<unsigned long long *>(ws.ch[0])[s * ws.stride] = bits;
}

// ---------------------------------------------------------------------------------- combine
template <bool IS_FLOAT>
__global__ void group_combine_kernel(GroupWs acc, GroupWs oth, int op, int64_t slots) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots) return;
    using Acc = typename std::conditional<IS_FLOAT, double, long long>::type;
    const int64_t w = s * acc.stride;
    Acc *a0 = reinterpret_cast<Acc *>(acc.ch[0]) + w, *a1 = reinterpret_cast<Acc *>(acc.ch[1]) + w;
    const Acc *o0 = reinterpret_cast<const Acc *>(oth.ch[0]) + w, *o1 = reinterpret_cast<const Acc *>(oth.ch[1]) + w;
    long long *ac = reinterpret_cast<long long *>(acc.ch[2]) + w;
    const long long *oc = reinterpret_cast<const long long *>(oth.ch[2]) + w;
    unsigned long long *ak = reinterpret_cast<unsigned long long *>(acc.ch[0]) + w;
    const unsigned long long *ok = reinterpret_cast<const unsigned long long *>(oth.ch[0]) + w;
    long long *ai = reinterpret_cast<long long *>(acc.ch[1]) + w;
    const long long *oi = reinterpret_cast<const long long *>(oth.ch[1]) + w;
    switch (op) {
        case NBG_GROUP_NANSUM:
        case NBG_GROUP_NANSUM_OF_SQUARES:
            *a0 = *a0 + *o0;
            break;
        case NBG_GROUP_NANCOUNT:
            *ac = *ac + *oc;
            break;
        case NBG_GROUP_NANMEAN:
            *a0 = *a0 + *o0;
            *ac = *ac + *oc;
            break;
        case NBG_GROUP_NANVAR:
        case NBG_GROUP_NANSTD:
            *a0 = *a0 + *o0;
            *a1 = *a1 + *o1;
            *ac = *ac + *oc;
            break;
        case NBG_GROUP_NANPROD:
            *a0 = *a0 * *o0;
            break;
        case NBG_GROUP_NANMIN:
        case NBG_GROUP_NANMAX:
            if (*ok > *ak) *ak = *ok;
            break;
        case NBG_GROUP_NANANY:
            if (*ok) *ak = 1;
            break;
        case NBG_GROUP_NANALL:
            if (!*ok) *ak = 0;
            break;
        case NBG_GROUP_NANFIRST:
            if (*oi < *ai) {
                *ai = *oi;
                *ak = *ok;
            }
            break;
        case NBG_GROUP_NANLAST:
            if (*oi > *ai) {
                *ai = *oi;
                *ak = *ok;
            }
            break;
        case NBG_GROUP_NANARGMAX:
        case NBG_GROUP_NANARGMIN:
            if (*ok > *ak) {
                *ak = *ok;
                *ai = *oi;
            } else if (*ok == *ak && *oi < *ai) {
                *ai = *oi;
            }
            break;
        default:
            break;
    }
}

// --------------------------------------------------------------------------------- finalize
template <typename V>
__global__ void group_finalize_kernel(GroupWs ws, V *__restrict__ out, int op, int64_t slots, int64_t ddof) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots) return;
    using Acc = typename VTraits<V>::Acc;
    const Acc a0 = reinterpret_cast<const Acc *>(ws.ch[0])[s * ws.stride];
    const Acc a1 = reinterpret_cast<const Acc *>(ws.ch[1])[s * ws.stride];
    const long long cnt = reinterpret_cast<const long long *>(ws.ch[2])[s * ws.stride];
    const unsigned long long k0 = reinterpret_cast<const unsigned long long *>(ws.ch[0])[s * ws.stride];
    const long long i1 = reinterpret_cast<const long long *>(ws.ch[1])[s * ws.stride];
    V r = (V)0;
    switch (op) {
        case NBG_GROUP_NANSUM:
        case NBG_GROUP_NANSUM_OF_SQUARES:
        case NBG_GROUP_NANPROD:
            r = (V)a0;
            break;
        case NBG_GROUP_NANCOUNT:
            r = (V)cnt;
            break;
        case NBG_GROUP_NANANY:
        case NBG_GROUP_NANALL:
            r = (V)(long long)k0;
            break;
        case NBG_GROUP_NANMEAN:
            // out[label] /= count: the V-typed sum divided in double (grouped.py:22-27)
            r = cnt == 0 ? VTraits<V>::nan_out() : (V)((double)(V)a0 / (double)cnt);
            break;
        case NBG_GROUP_NANVAR:
        case NBG_GROUP_NANSTD: {
            const long long denom = cnt - ddof;
            if (denom <= 0) {
                r = VTraits<V>::nan_out();
            } else {
                // (sums_of_squares - sums**2 / count) / denom with V-typed sums (grouped.py:176)
                const V sv = (V)a0, ssv = (V)a1;
                const double s2 = (double)sq_as_input(sv);
                const double num = __dsub_rn((double)ssv, s2 / (double)cnt);
                const double q = num / (double)denom;
                r = (V)(op == NBG_GROUP_NANSTD ? sqrt(q) : q);
            }
            break;
        }
        case NBG_GROUP_NANMAX:
            r = k0 == 0 ? VTraits<V>::nan_out() : (V)key_to_double(k0);
            break;
        case NBG_GROUP_NANMIN:
            r = k0 == 0 ? VTraits<V>::nan_out() : (V)key_to_double(~k0);
            break;
        case NBG_GROUP_NANFIRST:
        case NBG_GROUP_NANLAST: {
            const bool none = (op == NBG_GROUP_NANFIRST) ? (i1 == kIdxNone) : (i1 < 0);
            if (none) {
                // grouped.py:109-110 / :114: NaN for floats; first leaves integers untouched (0 here)
                r = (op == NBG_GROUP_NANFIRST && !VTraits<V>::is_float) ? (V)0 : VTraits<V>::nan_out();
            } else if (sizeof(V) == 8) {
                r = *reinterpret_cast<const V *>(&k0);
            } else {
                const unsigned int b = (unsigned int)k0;
                r = *reinterpret_cast<const V *>(&b);
            }
            break;
        }
        case NBG_GROUP_NANARGMAX:
        case NBG_GROUP_NANARGMIN:
            r = k0 == 0 ? VTraits<V>::nan_out() : (V)i1;
            break;
        default:
            break;
    }
    out[s] = r;
}

// ------------------------------------------------------------------------- L2 residency
// The per-element-label path resolves one atomic per element in L2; the accumulator table must
// stay there while tens of GB of inputs stream past it.  The table is marked PERSISTING (it may
// use the L2 set-aside) and everything else the kernel touches STREAMING, through the stream's
// access-policy window; the window is cleared again after the launch (stream-ordered).
struct L2Window {
    cudaStream_t st;
    bool on;
};
static L2Window l2_window_begin(cudaStream_t st, void *base, size_t bytes) {
    L2Window w{st, false};
    const char *e = getenv("NBG_L2_PERSIST");
    if (!e || atoi(e) == 0) return w;  // opt-in: the set-aside it needs halves the scan kernels' throughput
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return w;
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0 || bytes < ((size_t)1 << 20)) return w;
    static std::atomic<unsigned> limit_set{0};
    if (!(limit_set.load() & (1u << dev))) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
        limit_set.fetch_or(1u << dev);
    }
    cudaStreamAttrValue attr = {};
    const size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    attr.accessPolicyWindow.base_ptr = base;
    attr.accessPolicyWindow.num_bytes = win;
    double ratio = (double)max_persist / (double)win;
    attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) {
        cudaGetLastError();
        return w;
    }
    w.on = true;
    return w;
}
static void l2_window_end(const L2Window &w) {
    if (!w.on) return;
    cudaStreamAttrValue attr = {};
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(w.st, cudaStreamAttributeAccessPolicyWindow, &attr);
}

// ---------------------------------------------------------------------------------- launch
static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

template <typename V, typename L, int OP>
static int launch_atomic(const V *values, const L *labels, int labels_per_row, GroupWs ws, int64_t rows, int64_t n,
                         int64_t K, int64_t index_offset, cudaStream_t stream) {
    const int64_t bpr = (n + 2047) / 2048;
    if (bpr * rows > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_group: grid too large");
    const unsigned grid = (unsigned)(bpr * rows);
    const size_t plane_bytes = (size_t)rows * (size_t)K * 8;
    const WsLayout lay = ws_layout(OP, rows * K);
    // evict_last hints on the table reductions: config 5 nansum 22.9 -> 12.6 ms, nanfirst 49.7 -> 11.6 ms,
    // no device-wide state (NBG_L2_HINT=0 switches them off for A/B runs)
    int l2_hint = 1;
    if (const char *e = getenv("NBG_L2_HINT")) l2_hint = atoi(e);
    // one pass with the L2 window `win` over the table bytes that pass hammers with atomics
    auto pass = [&](auto kern, void *win, size_t win_bytes, const char *what) -> int {
        const L2Window l2 = l2_window_begin(stream, win, win_bytes);
        kern<<<grid, 256, 0, stream>>>(values, labels, labels_per_row, ws, rows, n, K, index_offset, bpr, l2_hint);
        const int rc = check_launch(what);
        l2_window_end(l2);
        return rc;
    };
    int rc;
    constexpr bool kMean = OP == NBG_GROUP_NANMEAN, kVar = OP == NBG_GROUP_NANVAR || OP == NBG_GROUP_NANSTD;
    if ((kMean || kVar) && lay.planar) {
        // table too large for L2.  One row of floating-point values: partition by label range, shared-memory
        // bins per range (nbg_group_partition.cuh) -- no global atomics on the table
        if constexpr (VTraits<V>::is_float) {
            if (rows == 1) {
                bool handled = false;
                rc = launch_partition<V, L, kVar ? 2 : 1>(values, labels, n, K, reinterpret_cast<double *>(ws.ch[0]),
                                                         reinterpret_cast<double *>(ws.ch[1]),
                                                         reinterpret_cast<long long *>(ws.ch[2]), stream, &handled);
                if (rc || handled) return rc;
            }
        }
        // otherwise: one pass per channel plane, that plane pinned
        rc = pass(group_atomic_kernel<V, L, OP, 0, 1>, ws.ch[0], plane_bytes, "nbg_group(atomic, sum plane)");
        if (rc) return rc;
        if (kVar) {
            rc = pass(group_atomic_kernel<V, L, OP, 0, 2>, ws.ch[1], plane_bytes, "nbg_group(atomic, sum-of-squares plane)");
            if (rc) return rc;
        }
        return pass(group_atomic_kernel<V, L, OP, 0, 4>, ws.ch[2], plane_bytes, "nbg_group(atomic, count plane)");
    }
    constexpr bool kArg = OP == NBG_GROUP_NANARGMAX || OP == NBG_GROUP_NANARGMIN;
    constexpr bool kEdge = OP == NBG_GROUP_NANFIRST || OP == NBG_GROUP_NANLAST;
    // arg*: both phases hammer / probe the key plane; first / last: the index plane; others: the table
    void *win = kEdge ? ws.ch[1] : ws.ch[0];
    if (OP == NBG_GROUP_NANCOUNT) win = ws.ch[2];
    const size_t win_bytes = (kArg || kEdge) ? plane_bytes : (size_t)lay.words * plane_bytes;
    rc = pass(group_atomic_kernel<V, L, OP, 0>, win, win_bytes, "nbg_group(atomic)");
    if (rc) return rc;
    if (kArg) {
        rc = pass(group_atomic_kernel<V, L, OP, 1>, win, win_bytes, "nbg_group(atomic, index phase)");
        if (rc) return rc;
    }
    if (kEdge) {
        group_gather_kernel<V><<<blocks_for(rows * K, 256), 256, 0, stream>>>(values, ws, rows, n, K, index_offset);
        rc = check_launch("nbg_group(gather)");
    }
    return rc;
}

template <typename V, typename L>
static int try_rowbins(int op, const V *values, const L *labels, GroupWs ws, void *scratch, size_t scratch_bytes,
                       int64_t rows, int64_t n, int64_t K, int64_t index_offset, cudaStream_t stream, bool *handled) {
    *handled = false;
    // conflict-free class-private bins first (one- to three-channel additive ops, prod, any/all)
#define NBG_RB2_CASE(CLS)                                                                                              \
    case CLS: {                                                                                                        \
        int rc2 = rb2_launch<V, L, CLS>(values, labels, ws.ch, ws.stride, scratch, scratch_bytes, rows, n, K, index_offset, stream, handled); \
        if (rc2 || *handled) return rc2;                                                                               \
        break;                                                                                                         \
    }
    switch (rb_class_of(op)) {
        NBG_RB2_CASE(RB_SUM)
        NBG_RB2_CASE(RB_COUNT)
        NBG_RB2_CASE(RB_MEAN)
        NBG_RB2_CASE(RB_SUMSQ)
        NBG_RB2_CASE(RB_VAR)
        NBG_RB2_CASE(RB_PROD)
        NBG_RB2_CASE(RB_ANY)
        NBG_RB2_CASE(RB_ALL)
        default:
            break;
    }
#undef NBG_RB2_CASE
#define NBG_RB_CASE(CLS) \
    case CLS:            \
        return rb_launch<V, L, CLS>(values, labels, ws.ch, ws.stride, scratch, scratch_bytes, rows, n, K, index_offset, stream, handled)
    switch (rb_class_of(op)) {
        NBG_RB_CASE(RB_SUM);
        NBG_RB_CASE(RB_COUNT);
        NBG_RB_CASE(RB_MEAN);
        NBG_RB_CASE(RB_SUMSQ);
        NBG_RB_CASE(RB_VAR);
        NBG_RB_CASE(RB_PROD);
        NBG_RB_CASE(RB_MAX);
        NBG_RB_CASE(RB_MIN);
        NBG_RB_CASE(RB_ARGMAX);
        NBG_RB_CASE(RB_ARGMIN);
        NBG_RB_CASE(RB_FIRST);
        NBG_RB_CASE(RB_LAST);
        NBG_RB_CASE(RB_ANY);
        NBG_RB_CASE(RB_ALL);
        default:
            return NBG_OK;
    }
#undef NBG_RB_CASE
}

template <typename V, typename L>
static int dispatch_accumulate(int op, const V *values, const L *labels, int labels_per_row, GroupWs ws, int64_t rows,
                               int64_t n, int64_t K, int64_t index_offset, void *scratch, size_t scratch_bytes,
                               cudaStream_t stream) {
    if (!labels_per_row) {
        bool handled = false;
        int rc = try_rowbins<V, L>(op, values, labels, ws, scratch, scratch_bytes, rows, n, K, index_offset, stream, &handled);
        if (rc || handled) return rc;
    }
#define NBG_GROUP_CASE(OPC) \
    case OPC:               \
        return launch_atomic<V, L, OPC>(values, labels, labels_per_row, ws, rows, n, K, index_offset, stream)
    switch (op) {
        NBG_GROUP_CASE(NBG_GROUP_NANMEAN);
        NBG_GROUP_CASE(NBG_GROUP_NANSUM);
        NBG_GROUP_CASE(NBG_GROUP_NANCOUNT);
        NBG_GROUP_CASE(NBG_GROUP_NANARGMAX);
        NBG_GROUP_CASE(NBG_GROUP_NANARGMIN);
        NBG_GROUP_CASE(NBG_GROUP_NANFIRST);
        NBG_GROUP_CASE(NBG_GROUP_NANLAST);
        NBG_GROUP_CASE(NBG_GROUP_NANPROD);
        NBG_GROUP_CASE(NBG_GROUP_NANSUM_OF_SQUARES);
        NBG_GROUP_CASE(NBG_GROUP_NANVAR);
        NBG_GROUP_CASE(NBG_GROUP_NANSTD);
        NBG_GROUP_CASE(NBG_GROUP_NANMIN);
        NBG_GROUP_CASE(NBG_GROUP_NANMAX);
        NBG_GROUP_CASE(NBG_GROUP_NANANY);
        NBG_GROUP_CASE(NBG_GROUP_NANALL);
        default:
            return fail(NBG_ERR_BAD_OP, "nbg_group: unknown op");
    }
#undef NBG_GROUP_CASE
}

template <typename V>
static int dispatch_labels(int op, int ldtype, const void *values, const void *labels, int labels_per_row, GroupWs ws,
                           int64_t rows, int64_t n, int64_t K, int64_t index_offset, void *scratch,
                           size_t scratch_bytes, cudaStream_t stream) {
    if (ldtype == NBG_I32)
        return dispatch_accumulate<V, int32_t>(op, static_cast<const V *>(values), static_cast<const int32_t *>(labels),
                                               labels_per_row, ws, rows, n, K, index_offset, scratch, scratch_bytes, stream);
    if (ldtype == NBG_I64)
        return dispatch_accumulate<V, int64_t>(op, static_cast<const V *>(values), static_cast<const int64_t *>(labels),
                                               labels_per_row, ws, rows, n, K, index_offset, scratch, scratch_bytes, stream);
    return fail(NBG_ERR_BAD_DTYPE, "nbg_group: labels dtype must be NBG_I32 or NBG_I64");
}

static bool float_only(int op) {
    return op == NBG_GROUP_NANMEAN || op == NBG_GROUP_NANVAR || op == NBG_GROUP_NANSTD;
}

}  // namespace nbg

static size_t group_state_bytes(int op, int64_t rows, int64_t num_labels) {
    return (size_t)nbg::ws_layout(op, rows * num_labels).words * (size_t)rows * (size_t)num_labels * 8 + 256;
}
extern "C" int nbg_group_record_words(int op) { return nbg::ws_layout(op).words; }
extern "C" int nbg_group_record_layout(int op, int64_t rows, int64_t num_labels, int64_t layout[5]) {
    using namespace nbg;
    if (op < 0 || op > NBG_GROUP_NANALL) return fail(NBG_ERR_BAD_OP, "nbg_group_record_layout: unknown op");
    if (!layout) return fail(NBG_ERR_BAD_ARG, "nbg_group_record_layout: null output");
    const WsLayout l = ws_layout(op, rows * num_labels);
    const int64_t plane = l.planar ? rows * num_labels : 1;
    layout[0] = l.stride;
    for (int c = 0; c < 3; c++) layout[1 + c] = (int64_t)l.slot[c] * plane;
    layout[4] = l.planar;
    return NBG_OK;
}
// scratch for the shared-label plan: 4 bytes per column + one header per (smallest) tile
static size_t group_scratch_bytes(int64_t n, int64_t rows) {
    // widest tile is 1536 columns (a short row still needs one whole tile), narrowest 128;
    // plus one lock word per group of 8 rows
    const size_t v1 = (size_t)(n + 2048) * 4 + (size_t)(n / 128 + 4) * nbg::kRbHdr * 4 + (size_t)(rows / 8 + 2) * 4 + 2048;
    const size_t v2 = nbg::rb2_scratch_bytes(n);
    return v1 > v2 ? v1 : v2;
}

extern "C" size_t nbg_group_workspace_bytes(int op, int, int64_t rows, int64_t n, int64_t num_labels) {
    if (rows <= 0 || num_labels <= 0) return 0;
    return group_state_bytes(op, rows, num_labels) + group_scratch_bytes(n > 0 ? n : 0, rows > 0 ? rows : 0);
}

static void *align256(void *p) { return reinterpret_cast<void *>(((uintptr_t)p + 255) & ~(uintptr_t)255); }

extern "C" int nbg_group_init(int op, int vdtype, void *workspace, int64_t rows, int64_t num_labels, void *stream) {
    using namespace nbg;
    if (op < 0 || op > NBG_GROUP_NANALL) return fail(NBG_ERR_BAD_OP, "nbg_group_init: unknown op");
    const int64_t slots = rows * num_labels;
    if (slots <= 0) return NBG_OK;
    if (!workspace) return fail(NBG_ERR_WORKSPACE, "nbg_group_init: null workspace");
    GroupWs ws = GroupWs::carve(align256(workspace), op, rows, num_labels);
    const int is_float = (vdtype == NBG_F32 || vdtype == NBG_F64);
    group_init_kernel<<<blocks_for(slots, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ws, op, is_float, slots);
    return check_launch("nbg_group_init");
}

extern "C" int nbg_group_accumulate(int op, int vdtype, int ldtype, const void *values, const void *labels,
                                    int labels_per_row, void *workspace, size_t workspace_bytes, int64_t rows,
                                    int64_t n, int64_t num_labels, int64_t index_offset, void *stream) {
    using namespace nbg;
    if (rows < 0 || n < 0 || num_labels < 0) return fail(NBG_ERR_BAD_ARG, "nbg_group: negative size");
    if (rows * n == 0 || num_labels == 0) return NBG_OK;
    if (!values || !labels || !workspace) return fail(NBG_ERR_BAD_ARG, "nbg_group: null pointer");
    if (float_only(op) && !(vdtype == NBG_F32 || vdtype == NBG_F64))
        return fail(NBG_ERR_BAD_DTYPE, "nbg_group: nanmean/nanvar/nanstd need float values (cast integers to float64)");
    if (workspace_bytes < group_state_bytes(op, rows, num_labels)) return fail(NBG_ERR_WORKSPACE, "nbg_group: workspace too small");
    GroupWs ws = GroupWs::carve(align256(workspace), op, rows, num_labels);
    // whatever follows the accumulator state is per-call scratch (column plan)
    unsigned char *scratch = static_cast<unsigned char *>(workspace) + group_state_bytes(op, rows, num_labels);
    const size_t scratch_bytes = workspace_bytes - group_state_bytes(op, rows, num_labels);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (vdtype) {
        case NBG_F32:
            return dispatch_labels<float>(op, ldtype, values, labels, labels_per_row, ws, rows, n, num_labels, index_offset, scratch, scratch_bytes, st);
        case NBG_F64:
            return dispatch_labels<double>(op, ldtype, values, labels, labels_per_row, ws, rows, n, num_labels, index_offset, scratch, scratch_bytes, st);
        case NBG_I32:
            return dispatch_labels<int32_t>(op, ldtype, values, labels, labels_per_row, ws, rows, n, num_labels, index_offset, scratch, scratch_bytes, st);
        case NBG_I64:
            return dispatch_labels<int64_t>(op, ldtype, values, labels, labels_per_row, ws, rows, n, num_labels, index_offset, scratch, scratch_bytes, st);
        default:
            return fail(NBG_ERR_BAD_DTYPE, "nbg_group: bad values dtype");
    }
}

extern "C" int nbg_group_combine(int op, int vdtype, void *accum, const void *other, int64_t rows, int64_t num_labels,
                                 void *stream) {
    using namespace nbg;
    const int64_t slots = rows * num_labels;
    if (slots <= 0) return NBG_OK;
    if (!accum || !other) return fail(NBG_ERR_BAD_ARG, "nbg_group_combine: null workspace");
    GroupWs a = GroupWs::carve(align256(accum), op, rows, num_labels);
    GroupWs o = GroupWs::carve(align256(const_cast<void *>(other)), op, rows, num_labels);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vdtype == NBG_F32 || vdtype == NBG_F64)
        group_combine_kernel<true><<<blocks_for(slots, 256), 256, 0, st>>>(a, o, op, slots);
    else
        group_combine_kernel<false><<<blocks_for(slots, 256), 256, 0, st>>>(a, o, op, slots);
    return check_launch("nbg_group_combine");
}

extern "C" int nbg_group_finalize(int op, int vdtype, const void *workspace, void *out, int64_t rows,
                                  int64_t num_labels, int64_t ddof, void *stream) {
    using namespace nbg;
    const int64_t slots = rows * num_labels;
    if (slots <= 0) return NBG_OK;
    if (!workspace || !out) return fail(NBG_ERR_BAD_ARG, "nbg_group_finalize: null pointer");
    GroupWs ws = GroupWs::carve(align256(const_cast<void *>(workspace)), op, rows, num_labels);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned nb = blocks_for(slots, 256);
    switch (vdtype) {
        case NBG_F32:
            group_finalize_kernel<float><<<nb, 256, 0, st>>>(ws, static_cast<float *>(out), op, slots, ddof);
            break;
        case NBG_F64:
            group_finalize_kernel<double><<<nb, 256, 0, st>>>(ws, static_cast<double *>(out), op, slots, ddof);
            break;
        case NBG_I32:
            group_finalize_kernel<int32_t><<<nb, 256, 0, st>>>(ws, static_cast<int32_t *>(out), op, slots, ddof);
            break;
        case NBG_I64:
            group_finalize_kernel<int64_t><<<nb, 256, 0, st>>>(ws, static_cast<int64_t *>(out), op, slots, ddof);
            break;
        default:
            return fail(NBG_ERR_BAD_DTYPE, "nbg_group_finalize: bad values dtype");
    }
    return check_launch("nbg_group_finalize");
}

extern "C" int nbg_group(int op, int vdtype, int ldtype, const void *values, const void *labels, int labels_per_row,
                         void *out, int64_t rows, int64_t n, int64_t num_labels, int64_t ddof, void *workspace,
                         size_t workspace_bytes, void *stream) {
    using namespace nbg;
    if (rows * num_labels > 0 && workspace_bytes < group_state_bytes(op, rows, num_labels))
        return fail(NBG_ERR_WORKSPACE, "nbg_group: workspace too small");
    int rc = nbg_group_init(op, vdtype, workspace, rows, num_labels, stream);
    if (rc) return rc;
    rc = nbg_group_accumulate(op, vdtype, ldtype, values, labels, labels_per_row, workspace, workspace_bytes, rows, n,
                              num_labels, 0, stream);
    if (rc) return rc;
    return nbg_group_finalize(op, vdtype, workspace, out, rows, num_labels, ddof, stream);
}
