// nbg_group_rowbins.cuh -- grouped sums when ONE label vector is shared by every row
// (groupndreduce with axis=int, numbagg/decorators.py:633-645: the xarray/flox case and
// BASELINE config 2).
//
// Idea: because all rows share the labels, the scatter pattern can be resolved ONCE per call
// and then replayed for every row without atomics.
//   plan kernel (labels only, O(n)): for every tile of C columns, the valid columns stably
//     sorted by label, packed as (label << 16 | column), plus 65 boundaries that cut the
//     sorted list into 64 label-aligned, balanced ranges.
//   main kernel: a CTA owns 8 rows x a column segment.  It streams row tiles (and the tile's
//     plan) through a 2-stage TMA ring into shared memory.  Its 512 threads form 64 sub-warps
//     of 8 lanes: lane = row, sub-warp = one of the 64 label ranges.  A sub-warp walks its
//     range; every entry is "bins[label][row] (+)= tile[row][column]" on a shared-memory bin
//     that no other lane can touch during this tile (ranges are label-aligned), so the
//     update is a plain load/add/store, and every bin receives its elements in ascending
//     column order -- the reference's own summation order (grouped.py:31-40), which makes the
//     float32 results bit-identical to numbagg's when a CTA covers whole rows.
//   Few labels (K * words <= 32, additive ops): ranges are cut evenly instead and each
//     sub-warp gets private bins that are summed at the end.  (Tried for the (value, index)
//     ops too: slower -- the private bins of a warp's four sub-warps share banks.)
// Bins are V-typed (accumulation in the OUTPUT dtype, like the reference) and are flushed
// into the common 8-byte workspace at the end (plain stores for one segment per row,
// atomics otherwise).
#pragma once

#include <stdlib.h>
#include <string.h>

#include "nbg_common.cuh"

namespace nbg {

constexpr int kRbRows = 8;    // rows per CTA = lanes per sub-warp
constexpr int kRbSub = 64;    // sub-warps per CTA
constexpr int kRbThreads = kRbRows * kRbSub;  // 512
constexpr int kRbHdr = 68;    // u32 words per tile header: bounds[65], count, 2 pad (272 B)

// Bin slot of a label.  The 64 label ranges a tile is cut into are, for labels that are
// roughly uniform, close to the NOMINAL ranges [s*W, (s+1)*W), W = ceil(K/64).  Slots are laid
// out so that the four sub-warps of a warp (nominal ranges 4g..4g+3) own slots = 0,1,2,3 mod 4:
// bins[slot][row] is 8 consecutive words per slot, so those four sub-warps then touch four
// different bank octets and their bin accesses stop conflicting.  (A pure performance
// permutation: any label distribution stays correct.)  W == 0: identity.
__host__ __device__ inline int rb_slot_of(int label, int W) {
    if (W == 0) return label;
    const int s_nom = label / W, j = label - s_nom * W;
    return ((s_nom >> 2) * W + j) * 4 + (s_nom & 3);
}
__host__ __device__ inline int rb_label_of(int slot, int W) {
    if (W == 0) return slot;
    const int q = slot & 3, t = slot >> 2;
    const int g = t / W, j = t - g * W;
    return (g * 4 + q) * W + j;
}
__host__ __device__ inline int rb_num_slots(int K, int W) { return W == 0 ? K : 64 * W; }

enum RbClass {
    RB_SUM = 0, RB_COUNT = 1, RB_MEAN = 2, RB_SUMSQ = 3, RB_VAR = 4,  // additive (mergeable with atomics)
    RB_PROD = 5, RB_MAX = 6, RB_MIN = 7, RB_ARGMAX = 8, RB_ARGMIN = 9, RB_FIRST = 10, RB_LAST = 11, RB_ANY = 12, RB_ALL = 13
};

// ----------------------------------------------------------------------------------- plan
template <typename L>
__global__ void __launch_bounds__(256) group_plan_kernel(const L *__restrict__ labels, int64_t n, int K, int C,
                                                         int even_split, int W, uint32_t *__restrict__ plan) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *offs = reinterpret_cast<int *>(smem_raw);        // [K + 1] histogram -> exclusive offsets
    int *cursor = offs + (K + 1);                          // [K]
    uint32_t *slab = reinterpret_cast<uint32_t *>(cursor + K);  // [C] label or 0xffffffff
    uint32_t *sent = slab + C;                             // [C] sorted entries
    __shared__ int warp_tot[8];
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * C;
    uint32_t *out = plan + (size_t)blockIdx.x * (kRbHdr + C);

    for (int k = tid; k <= K; k += 256) offs[k] = 0;
    __syncthreads();
    for (int j = tid; j < C; j += 256) {
        const int64_t col = c0 + j;
        long long lab = -1;
        if (col < n) lab = (long long)labels[col];
        const bool valid = lab >= 0 && lab < K;
        slab[j] = valid ? (uint32_t)lab : 0xffffffffu;
        sent[j] = 0;
        if (valid) atomicAdd(&offs[(int)lab], 1);
    }
    __syncthreads();
    // exclusive scan of offs[0..K) (block of 256 threads, contiguous runs per thread)
    {
        const int per = (K + 255) / 256;
        const int beg = min(tid * per, K), end = min(beg + per, K);
        int local = 0;
        for (int k = beg; k < end; k++) local += offs[k];
        int inc = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int base = 0;
        for (int w2 = 0; w2 < wid; w2++) base += warp_tot[w2];
        int run = base + inc - local;
        for (int k = beg; k < end; k++) {
            const int c = offs[k];
            offs[k] = run;
            cursor[k] = run;
            run += c;
        }
        if (tid == 255) {
            int tot = 0;
            for (int w2 = 0; w2 < 8; w2++) tot += warp_tot[w2];
            s_count = tot;
            offs[K] = tot;
        }
    }
    __syncthreads();
    // stable placement by one warp: columns in ascending order, 32 at a time
    if (wid == 0) {
        for (int j0 = 0; j0 < C; j0 += 32) {
            const int j = j0 + lane;
            const uint32_t lab = j < C ? slab[j] : 0xffffffffu;
            const bool valid = lab != 0xffffffffu;
            const unsigned m = __match_any_sync(0xffffffffu, lab);
            if (valid) {
                const int rank = __popc(m & ((1u << lane) - 1u));
                const int base = cursor[lab];
                sent[base + rank] = ((uint32_t)rb_slot_of((int)lab, W) << 16) | (uint32_t)j;
            }
            __syncwarp();
            if (valid && (m & ((1u << lane) - 1u)) == 0) cursor[lab] += __popc(m);
            __syncwarp();
        }
    }
    __syncthreads();
    const int count = s_count;
    if (tid <= kRbSub) {
        int b;
        if (tid == kRbSub || count == 0) {
            b = count;
        } else {
            const int e = (int)(((long long)tid * count) / kRbSub);
            b = even_split ? e : offs[rb_label_of((int)(sent[e] >> 16), W)];  // start of the label run containing e
        }
        out[tid] = (uint32_t)b;
    }
    if (tid == kRbSub + 1) out[kRbSub + 1] = (uint32_t)count;
    if (tid == kRbSub + 2) out[kRbSub + 2] = 0;
    if (tid == kRbSub + 3) out[kRbSub + 3] = 0;
    for (int j = tid; j < C; j += 256) out[kRbHdr + j] = sent[j];
}

// ----------------------------------------------------------------------------------- bins
template <typename V>
struct RbCounter {
    using type = int32_t;
};
template <>
struct RbCounter<double> {
    using type = long long;
};
template <>
struct RbCounter<int64_t> {
    using type = long long;
};

template <typename V>
__device__ __forceinline__ V v_add(V a, V b) {
    return a + b;
}
template <>
__device__ __forceinline__ float v_add<float>(float a, float b) {
    return __fadd_rn(a, b);
}
template <>
__device__ __forceinline__ double v_add<double>(double a, double b) {
    return __dadd_rn(a, b);
}
template <typename V>
__device__ __forceinline__ V v_sq(V a) {
    return (V)((unsigned long long)a * (unsigned long long)a);
}
template <>
__device__ __forceinline__ float v_sq<float>(float a) {
    return __fmul_rn(a, a);
}
template <>
__device__ __forceinline__ double v_sq<double>(double a) {
    return __dmul_rn(a, a);
}
template <>
__device__ __forceinline__ int32_t v_sq<int32_t>(int32_t a) {
    return (int32_t)((uint32_t)a * (uint32_t)a);
}

template <typename V, int CLS>
struct RbBin;
template <typename V>
struct RbBin<V, RB_SUM> {
    V s;
    __device__ __forceinline__ void zero() { s = (V)0; }
    // `ok` false (NaN observation): adds +0, which leaves every reachable sum unchanged
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) { s = v_add(s, ok ? v : (V)0); }
    __device__ __forceinline__ void merge(const RbBin &o) { s = v_add(s, o.s); }
};
template <typename V>
struct RbBin<V, RB_SUMSQ> {
    V s;
    __device__ __forceinline__ void zero() { s = (V)0; }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) { s = v_add(s, v_sq(ok ? v : (V)0)); }
    __device__ __forceinline__ void merge(const RbBin &o) { s = v_add(s, o.s); }
};
template <typename V>
struct RbBin<V, RB_COUNT> {
    typename RbCounter<V>::type c;
    __device__ __forceinline__ void zero() { c = 0; }
    template <typename I>
    __device__ __forceinline__ void add(V, bool ok, I) { c += ok ? 1 : 0; }
    __device__ __forceinline__ void merge(const RbBin &o) { c += o.c; }
};
template <typename V>
struct alignas(2 * sizeof(V)) RbBin<V, RB_MEAN> {
    V s;
    typename RbCounter<V>::type c;
    __device__ __forceinline__ void zero() {
        s = (V)0;
        c = 0;
    }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) {
        s = v_add(s, ok ? v : (V)0);
        c += ok ? 1 : 0;
    }
    __device__ __forceinline__ void merge(const RbBin &o) {
        s = v_add(s, o.s);
        c += o.c;
    }
};
template <typename V>
struct alignas(4 * sizeof(V)) RbBin<V, RB_VAR> {
    V s, ss;
    typename RbCounter<V>::type c, pad;
    __device__ __forceinline__ void zero() {
        s = (V)0;
        ss = (V)0;
        c = 0;
        pad = 0;
    }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) {
        const V m = ok ? v : (V)0;
        s = v_add(s, m);
        ss = v_add(ss, v_sq(m));
        c += ok ? 1 : 0;
    }
    __device__ __forceinline__ void merge(const RbBin &o) {
        s = v_add(s, o.s);
        ss = v_add(ss, o.ss);
        c += o.c;
    }
};

// ---- non-additive classes: the reference's loop bodies on a privately owned bin ----------
template <typename V>
__device__ __forceinline__ V v_mul(V a, V b) {
    return (V)((unsigned long long)a * (unsigned long long)b);
}
template <>
__device__ __forceinline__ float v_mul<float>(float a, float b) {
    return __fmul_rn(a, b);
}
template <>
__device__ __forceinline__ double v_mul<double>(double a, double b) {
    return __dmul_rn(a, b);
}
template <>
__device__ __forceinline__ int32_t v_mul<int32_t>(int32_t a, int32_t b) {
    return (int32_t)((uint32_t)a * (uint32_t)b);
}

template <typename V>
struct RbBin<V, RB_PROD> {  // grouped.py:124-132
    V pr;
    __device__ __forceinline__ void zero() { pr = (V)1; }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) { pr = v_mul(pr, ok ? v : (V)1); }
    __device__ __forceinline__ void merge(const RbBin &o) { pr = v_mul(pr, o.pr); }
};
template <typename V, bool IS_MAX>
struct alignas(2 * sizeof(V)) RbExtreme {  // grouped.py:211-244: compare as double, first extreme wins
    V best;
    typename RbCounter<V>::type has;
    __device__ __forceinline__ void zero() {
        best = (V)0;
        has = 0;
    }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) {
        const bool better = IS_MAX ? ((double)v > (double)best) : ((double)v < (double)best);
        const bool take = ok && (!has || better);
        best = take ? v : best;
        has |= ok ? 1 : 0;
    }
};
template <typename V>
struct RbBin<V, RB_MAX> : RbExtreme<V, true> {};
template <typename V>
struct RbBin<V, RB_MIN> : RbExtreme<V, false> {};
// (value, position) bins: two words, `idx < 0` = nothing taken yet (idx is the column within
// this shard: < 2^31 on this path, so it has the width of the value).
template <typename V, bool IS_MAX>
struct alignas(2 * sizeof(V)) RbArgExtreme {  // grouped.py:54-92
    V best;
    typename RbCounter<V>::type idx;
    __device__ __forceinline__ void zero() {
        best = (V)0;
        idx = -1;
    }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I gi) {
        const bool better = IS_MAX ? ((double)v > (double)best) : ((double)v < (double)best);
        const bool take = ok && (idx < 0 || better);
        best = take ? v : best;
        idx = take ? (typename RbCounter<V>::type)gi : idx;
    }
};
template <typename V>
struct RbBin<V, RB_ARGMAX> : RbArgExtreme<V, true> {};
template <typename V>
struct RbBin<V, RB_ARGMIN> : RbArgExtreme<V, false> {};
template <typename V, bool FIRST>
struct alignas(2 * sizeof(V)) RbEdge {  // grouped.py:95-121: first / last valid value (+ its index)
    V val;
    typename RbCounter<V>::type idx;
    __device__ __forceinline__ void zero() {
        val = (V)0;
        idx = -1;
    }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I gi) {
        const bool take = FIRST ? (ok && idx < 0) : ok;
        val = take ? v : val;
        idx = take ? (typename RbCounter<V>::type)gi : idx;
    }
};
template <typename V>
struct RbBin<V, RB_FIRST> : RbEdge<V, true> {};
template <typename V>
struct RbBin<V, RB_LAST> : RbEdge<V, false> {};
template <typename V>
struct RbBin<V, RB_ANY> {  // grouped.py:247-257
    typename RbCounter<V>::type flag;
    __device__ __forceinline__ void zero() { flag = 0; }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) { flag |= (ok && v != (V)0) ? 1 : 0; }
    __device__ __forceinline__ void merge(const RbBin &o) { flag |= o.flag; }
};
template <typename V>
struct RbBin<V, RB_ALL> {  // grouped.py:260-270
    typename RbCounter<V>::type flag;
    __device__ __forceinline__ void zero() { flag = 1; }
    template <typename I>
    __device__ __forceinline__ void add(V v, bool ok, I) { flag &= (ok && v == (V)0) ? 0 : 1; }
    __device__ __forceinline__ void merge(const RbBin &o) { flag &= o.flag; }
};

template <typename V, int CLS>
struct RbFlush;  // bin -> the three 8-byte workspace channels (Acc ch0, Acc ch1, i64 ch2)

struct RbParams {
    const void *values;
    const uint32_t *plan;
    void *ws_ch[3];
    int64_t ws_stride;
    int64_t rows, n;
    int K, C;
    int ntiles, tiles_per_seg, nseg;
    int priv;  // sub-warp-private bins (few labels)
    int W;     // nominal label-range width of the slot permutation (0: identity)
    int64_t index_offset;  // flat index of column 0 of this shard (arg* / first / last)
    int *locks;            // one per row group: serialises the record merges of (value, index) ops across column segments
};

template <typename V>
__host__ __device__ inline int rb_row_stride(int C) {
    return C + 16 / (int)sizeof(V);  // +16 bytes: rows start 4 banks apart
}
template <typename V, int CLS>
__host__ __device__ inline size_t rb_stage_bytes(int C) {
    return ((size_t)(kRbHdr + C) * 4 + (size_t)kRbRows * rb_row_stride<V>(C) * sizeof(V) + 15) & ~(size_t)15;
}
// the slot permutation only pays while one slot (8 rows x bin) spans fewer than 32 banks
template <typename V, int CLS>
__host__ __device__ inline int rb_nominal_width(int K, int priv) {
    return (priv || sizeof(RbBin<V, CLS>) > 8) ? 0 : (K + 63) / 64;
}
template <typename V, int CLS>
__host__ __device__ inline size_t rb_bins_bytes(int K, int priv) {
    const int slots = rb_num_slots(K, rb_nominal_width<V, CLS>(K, priv));
    return ((size_t)slots * kRbRows * (priv ? kRbSub : 1) * sizeof(RbBin<V, CLS>) + 15) & ~(size_t)15;
}
template <typename V, int CLS>
__host__ __device__ inline size_t rb_smem_bytes(int K, int C, int priv) {
    return 64 + rb_bins_bytes<V, CLS>(K, priv) + 2 * rb_stage_bytes<V, CLS>(C);
}

template <typename V, int CLS>
__global__ void __launch_bounds__(kRbThreads) group_rowbins_kernel(RbParams p) {
    using Bin = RbBin<V, CLS>;
    using Acc = typename std::conditional<std::is_floating_point<V>::value, double, long long>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);  // [2]
    Bin *bins = reinterpret_cast<Bin *>(smem_raw + 64);
    unsigned char *stage0 = smem_raw + 64 + rb_bins_bytes<V, CLS>(p.K, p.priv);
    const size_t stage_bytes = rb_stage_bytes<V, CLS>(p.C);
    const int C = p.C, K = p.K;
    const int stride = rb_row_stride<V>(C);

    const int tid = threadIdx.x;
    const int r = tid & (kRbRows - 1);   // row within the group
    const int sub = tid >> 3;            // sub-warp = label range
    const int64_t g = blockIdx.x / p.nseg;
    const int seg = blockIdx.x % p.nseg;
    const int64_t r0 = g * kRbRows;
    const int nrows = (int)min((int64_t)kRbRows, p.rows - r0);
    const int t_beg = seg * p.tiles_per_seg;
    const int t_end = min(t_beg + p.tiles_per_seg, p.ntiles);
    const V *vbase = reinterpret_cast<const V *>(p.values) + r0 * p.n;

    const int nslots = rb_num_slots(K, p.W);
    const int nbins = nslots * kRbRows * (p.priv ? kRbSub : 1);
    for (int i = tid; i < nbins; i += kRbThreads) bins[i].zero();
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t, int st) {
        // thread 0 only: plan block + one bulk copy per row of the tile
        unsigned char *sb = stage0 + (size_t)st * stage_bytes;
        const int64_t c0 = (int64_t)t * C;
        const int cols = (int)min((int64_t)C, p.n - c0);
        const uint32_t plan_bytes = (uint32_t)(kRbHdr + C) * 4u;
        const uint32_t row_bytes = (uint32_t)cols * (uint32_t)sizeof(V);
        mbar_arrive_expect_tx(&bar[st], plan_bytes + (uint32_t)nrows * row_bytes);
        bulk_g2s(sb, p.plan + (size_t)t * (kRbHdr + C), plan_bytes, &bar[st]);
        V *tile = reinterpret_cast<V *>(sb + (size_t)(kRbHdr + C) * 4);
        for (int rr = 0; rr < nrows; rr++)
            bulk_g2s(tile + (size_t)rr * stride, vbase + (int64_t)rr * p.n + c0, row_bytes, &bar[st]);
    };

    if (tid == 0 && t_beg < t_end) issue(t_beg, 0);
    // 32-bit shared-window addresses: the RMW stream below must compile to LDS/STS on
    // register+immediate addresses, not to generic loads
    const uint32_t bins_s = smem_u32(bins) + (uint32_t)((p.priv ? (size_t)sub * K * kRbRows : 0) + r) * (uint32_t)sizeof(Bin);
    constexpr uint32_t kLabelStep = kRbRows * (uint32_t)sizeof(Bin);  // bytes between labels
    auto bin_at = [&](uint32_t e) -> Bin * {
        return reinterpret_cast<Bin *>(__cvta_shared_to_generic(bins_s + (e >> 16) * kLabelStep));
    };
    for (int t = t_beg, it = 0; t < t_end; t++, it++) {
        const int st = it & 1;
        if (tid == 0 && t + 1 < t_end) issue(t + 1, st ^ 1);  // stage st^1 was drained last iteration
        mbar_wait(&bar[st], (uint32_t)((it >> 1) & 1));
        const uint32_t sb = smem_u32(stage0 + (size_t)st * stage_bytes);
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(__cvta_shared_to_generic(sb));
        const uint32_t ent_s = sb + kRbHdr * 4;
        const uint32_t tile_s = sb + (uint32_t)(kRbHdr + C) * 4 + (uint32_t)r * (uint32_t)stride * (uint32_t)sizeof(V);
        const int b0 = (int)hdr[sub], b1 = (int)hdr[sub + 1];
        const int cbase = t * C;  // column of tile entry 0 within the row (n < 2^31 on this path)
        auto ent_at = [&](int i) -> uint32_t {
            return *reinterpret_cast<const uint32_t *>(__cvta_shared_to_generic(ent_s + (uint32_t)i * 4u));
        };
        auto val_at = [&](uint32_t e) -> V {
            return *reinterpret_cast<const V *>(__cvta_shared_to_generic(tile_s + (e & 0xffffu) * (uint32_t)sizeof(V)));
        };
        if (r < nrows) {
            int i = b0;
            for (; i + 4 <= b1; i += 4) {
                const uint32_t e0 = ent_at(i), e1 = ent_at(i + 1), e2 = ent_at(i + 2), e3 = ent_at(i + 3);
                const V v0 = val_at(e0), v1 = val_at(e1), v2 = val_at(e2), v3 = val_at(e3);
                // read-modify-write in entry order: consecutive entries may share a label
                Bin *q0 = bin_at(e0);
                { Bin b = *q0; b.add(v0, !is_nan(v0), cbase + (int)(e0 & 0xffffu)); *q0 = b; }
                Bin *q1 = bin_at(e1);
                { Bin b = *q1; b.add(v1, !is_nan(v1), cbase + (int)(e1 & 0xffffu)); *q1 = b; }
                Bin *q2 = bin_at(e2);
                { Bin b = *q2; b.add(v2, !is_nan(v2), cbase + (int)(e2 & 0xffffu)); *q2 = b; }
                Bin *q3 = bin_at(e3);
                { Bin b = *q3; b.add(v3, !is_nan(v3), cbase + (int)(e3 & 0xffffu)); *q3 = b; }
            }
            for (; i < b1; i++) {
                const uint32_t e = ent_at(i);
                const V v = val_at(e);
                Bin *q = bin_at(e);
                Bin b = *q;
                b.add(v, !is_nan(v), cbase + (int)(e & 0xffffu));
                *q = b;
            }
        }
        __syncthreads();  // everyone is done with stage st before it is refilled
    }

    // ---- flush bins into the 8-byte workspace channels
    Acc *c0p = reinterpret_cast<Acc *>(p.ws_ch[0]);
    Acc *c1p = reinterpret_cast<Acc *>(p.ws_ch[1]);
    long long *c2p = reinterpret_cast<long long *>(p.ws_ch[2]);
    const bool atomic = p.nseg > 1;
    // (value, index) records are two words that must change together: the column segments of
    // a row group take turns (spin lock per group; the holder is a resident CTA that needs
    // nothing from the waiters, so this cannot deadlock).  The merge itself is order-free.
    constexpr bool kPair = CLS == RB_ARGMAX || CLS == RB_ARGMIN || CLS == RB_FIRST || CLS == RB_LAST;
    const bool locked = kPair && atomic;
    if (locked) {
        if (tid == 0) {
            while (atomicCAS(&p.locks[g], 0, 1) != 0) __nanosleep(100);
            __threadfence();
        }
        __syncthreads();
    }
    for (int idx = tid; idx < K * kRbRows; idx += kRbThreads) {
        const int rr = idx / K, k = idx - rr * K;  // consecutive threads -> consecutive labels
        if (rr >= nrows) continue;
        Bin b = bins[rb_slot_of(k, p.W) * kRbRows + rr];
        if constexpr (CLS <= RB_VAR) {
            if (p.priv) {
                for (int s2 = 1; s2 < kRbSub; s2++) b.merge(bins[(size_t)s2 * K * kRbRows + k * kRbRows + rr]);
            }
        }
        const size_t o = ((size_t)(r0 + rr) * K + k) * (size_t)p.ws_stride;  // record offset (words)
        if constexpr (CLS <= RB_VAR) {
            RbFlush<V, CLS>::flush(b, c0p + o, c1p + o, c2p + o, atomic);
        } else {
            RbFlush<V, CLS>::flush(b, c0p + o, c1p + o, p.index_offset, atomic);
        }
    }
    if (locked) {
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(&p.locks[g], 0);
    }
}

template <typename A>
__device__ __forceinline__ void rb_put(A *p, A v, bool atomic) {
    if (atomic) {
        if (sizeof(A) == 8 && std::is_floating_point<A>::value)
            atomicAdd(reinterpret_cast<double *>(p), (double)v);
        else
            atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
    } else {
        *p = *p + v;  // the workspace was initialised (and may hold earlier shards)
    }
}

template <typename V>
struct RbFlush<V, RB_SUM> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_SUM> &b, A *c0, A *, long long *, bool at) {
        rb_put(c0, (A)b.s, at);
    }
};
template <typename V>
struct RbFlush<V, RB_SUMSQ> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_SUMSQ> &b, A *c0, A *, long long *, bool at) {
        rb_put(c0, (A)b.s, at);
    }
};
template <typename V>
struct RbFlush<V, RB_COUNT> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_COUNT> &b, A *, A *, long long *c2, bool at) {
        rb_put(c2, (long long)b.c, at);
    }
};
template <typename V>
struct RbFlush<V, RB_MEAN> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_MEAN> &b, A *c0, A *, long long *c2, bool at) {
        rb_put(c0, (A)b.s, at);
        rb_put(c2, (long long)b.c, at);
    }
};
template <typename V>
struct RbFlush<V, RB_VAR> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_VAR> &b, A *c0, A *c1, long long *c2, bool at) {
        rb_put(c0, (A)b.s, at);
        rb_put(c1, (A)b.ss, at);
        rb_put(c2, (long long)b.c, at);
    }
};


// Non-additive classes: merge the bin into the workspace record (this CTA is the only writer of
// its rows' records; the record may already hold earlier shards of the same device).
template <typename V>
struct RbFlush<V, RB_PROD> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_PROD> &b, A *c0, A *, int64_t, bool atomic) {
        // V-typed product (keeps the float32 rounding of the reference), CAS loop when several
        // column segments share the record
        unsigned long long *w = reinterpret_cast<unsigned long long *>(c0);
        unsigned long long old = *w, assumed;
        do {
            assumed = old;
            A cur;
            memcpy(&cur, &assumed, 8);
            A next;
            if (std::is_floating_point<A>::value) next = (A)v_mul((V)cur, b.pr);
            else next = (A)((unsigned long long)cur * (unsigned long long)(long long)b.pr);
            unsigned long long nb;
            memcpy(&nb, &next, 8);
            if (!atomic) {
                *w = nb;
                break;
            }
            old = atomicCAS(w, assumed, nb);
        } while (old != assumed);
    }
};
template <typename V, int CLS>
struct RbFlushExtreme {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, CLS> &b, A *c0, A *, int64_t, bool atomic) {
        if (!b.has) return;
        unsigned long long k = order_key((double)b.best);
        if (CLS == RB_MIN) k = ~k;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(c0);
        if (atomic) atomicMax(w, k);
        else if (k > *w) *w = k;
    }
};
template <typename V>
struct RbFlush<V, RB_MAX> : RbFlushExtreme<V, RB_MAX> {};
template <typename V>
struct RbFlush<V, RB_MIN> : RbFlushExtreme<V, RB_MIN> {};
template <typename V, int CLS>
struct RbFlushArg {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, CLS> &b, A *c0, A *c1, int64_t off, bool) {
        if (b.idx < 0) return;
        unsigned long long k = order_key((double)b.best);
        if (CLS == RB_ARGMIN) k = ~k;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(c0);
        long long *wi = reinterpret_cast<long long *>(c1);
        const long long gi = (long long)b.idx + off;
        const unsigned long long cur = __ldcg(w);  // past L1: another SM's segment may have merged meanwhile
        if (k > cur) {
            *w = k;
            *wi = gi;
        } else if (k == cur && gi < __ldcg(wi)) {
            *wi = gi;
        }
    }
};
template <typename V>
struct RbFlush<V, RB_ARGMAX> : RbFlushArg<V, RB_ARGMAX> {};
template <typename V>
struct RbFlush<V, RB_ARGMIN> : RbFlushArg<V, RB_ARGMIN> {};
template <typename V, int CLS>
struct RbFlushEdge {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, CLS> &b, A *c0, A *c1, int64_t off, bool) {
        if (b.idx < 0) return;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(c0);
        long long *wi = reinterpret_cast<long long *>(c1);
        const long long gi = (long long)b.idx + off;
        const long long cur = __ldcg(wi);  // past L1, see RbFlushArg
        const bool take = CLS == RB_FIRST ? (gi < cur) : (gi > cur);
        if (take) {
            unsigned long long bits;
            if (sizeof(V) == 8) bits = *reinterpret_cast<const unsigned long long *>(&b.val);
            else bits = (unsigned long long)*reinterpret_cast<const unsigned int *>(&b.val);
            *w = bits;
            *wi = gi;
        }
    }
};
template <typename V>
struct RbFlush<V, RB_FIRST> : RbFlushEdge<V, RB_FIRST> {};
template <typename V>
struct RbFlush<V, RB_LAST> : RbFlushEdge<V, RB_LAST> {};
template <typename V>
struct RbFlush<V, RB_ANY> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_ANY> &b, A *c0, A *, int64_t, bool) {
        if (b.flag) *reinterpret_cast<long long *>(c0) = 1;
    }
};
template <typename V>
struct RbFlush<V, RB_ALL> {
    template <typename A>
    __device__ static __forceinline__ void flush(const RbBin<V, RB_ALL> &b, A *c0, A *, int64_t, bool) {
        if (!b.flag) *reinterpret_cast<long long *>(c0) = 0;
    }
};

// ------------------------------------------------------------------------------------ host
inline int rb_class_of(int op) {
    switch (op) {
        case NBG_GROUP_NANSUM:
            return RB_SUM;
        case NBG_GROUP_NANCOUNT:
            return RB_COUNT;
        case NBG_GROUP_NANMEAN:
            return RB_MEAN;
        case NBG_GROUP_NANSUM_OF_SQUARES:
            return RB_SUMSQ;
        case NBG_GROUP_NANVAR:
        case NBG_GROUP_NANSTD:
            return RB_VAR;
        case NBG_GROUP_NANPROD:
            return RB_PROD;
        case NBG_GROUP_NANMAX:
            return RB_MAX;
        case NBG_GROUP_NANMIN:
            return RB_MIN;
        case NBG_GROUP_NANARGMAX:
            return RB_ARGMAX;
        case NBG_GROUP_NANARGMIN:
            return RB_ARGMIN;
        case NBG_GROUP_NANFIRST:
            return RB_FIRST;
        case NBG_GROUP_NANLAST:
            return RB_LAST;
        case NBG_GROUP_NANANY:
            return RB_ANY;
        case NBG_GROUP_NANALL:
            return RB_ALL;
        default:
            return -1;
    }
}

struct RbGeometry {
    bool ok;
    int C, priv, ntiles, nseg, tiles_per_seg;
    size_t smem, plan_bytes, lock_bytes;
};

template <typename V, int CLS>
static RbGeometry rb_geometry(int64_t rows, int64_t n, int64_t K) {
    RbGeometry g = {};
    constexpr int PER16 = 16 / (int)sizeof(V);
    if (K <= 0 || K > 65535 || n <= 0 || n >= ((int64_t)1 << 31) || (n % PER16) != 0 || rows < 1) return g;
    const int words = (int)(sizeof(RbBin<V, CLS>) / sizeof(V));
    g.priv = (CLS <= RB_VAR && K * words <= 32) ? 1 : 0;
    // Tile width C (columns per stage) and CTAs per SM, in measured order of preference on
    // config 2 (group_nansum / nanmean / nanstd float32, 10^4 x 10^6, 1000 labels):
    // wide tiles amortise the per-tile barrier and plan header, two CTAs per SM hide the
    // other one's pipeline bubbles -- 4 KB rows x 2 CTAs (61 %) > 6 KB x 1 (56 % sum, 46 %
    // mean) > 2 KB x 2 (51 %, 41 %) > 4 KB x 1 (38 % mean) > ...
    const int kb = 1024 / (int)sizeof(V);  // elements per KB of row
    struct Cand {
        int C;
        bool two;
    };
    Cand order[7] = {{4 * kb, true}, {6 * kb, false}, {2 * kb, true}, {4 * kb, false},
                     {2 * kb, false}, {kb, true}, {kb, false}};
    if (const char *e = getenv("NBG_RB_C")) {  // tuning hook
        for (auto &c : order) c.C = atoi(e);
    }
    for (const auto &c : order) {
        const size_t s = rb_smem_bytes<V, CLS>((int)K, c.C, g.priv);
        if (s <= (c.two ? (size_t)110 * 1024 : kMaxSmem)) {
            g.C = c.C;
            g.smem = s;
            g.ok = true;
            break;
        }
    }
    if (!g.ok) return g;
    g.ntiles = (int)((n + g.C - 1) / g.C);
    const int64_t groups = (rows + kRbRows - 1) / kRbRows;
    // Whole rows per CTA (=> the reference's exact summation order) once there are >= 4 waves
    // of row groups; otherwise cut rows into column segments for ~8 waves and merge the
    // partial bins with atomics.
    const int64_t slots = (int64_t)kNumSMs * 2;
    int64_t nseg = groups >= 4 * slots ? 1 : (8 * slots + groups - 1) / groups;
    if (const char *e = getenv("NBG_RB_NSEG")) nseg = atoi(e);  // tuning hook
    if (nseg > g.ntiles) nseg = g.ntiles;
    if (nseg < 1) nseg = 1;
    g.tiles_per_seg = (int)((g.ntiles + nseg - 1) / nseg);
    g.nseg = (g.ntiles + g.tiles_per_seg - 1) / g.tiles_per_seg;
    g.plan_bytes = (size_t)g.ntiles * (kRbHdr + g.C) * 4 + 256;
    g.lock_bytes = (size_t)(groups + 1) * sizeof(int);
    return g;
}

template <typename V, typename L, int CLS>
static int rb_launch(const V *values, const L *labels, void *ws_ch[3], int64_t ws_stride, void *scratch, size_t scratch_bytes,
                     int64_t rows, int64_t n, int64_t K, int64_t index_offset, cudaStream_t stream, bool *handled) {
    *handled = false;
    const RbGeometry g = rb_geometry<V, CLS>(rows, n, K);
    if (!g.ok || scratch == nullptr || scratch_bytes < g.plan_bytes + g.lock_bytes + 256) return NBG_OK;
    if (((uintptr_t)values & 15) != 0) return NBG_OK;
    uint32_t *plan = reinterpret_cast<uint32_t *>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    const size_t plan_smem = (size_t)(2 * K + 1) * 4 + (size_t)2 * g.C * 4 + 16;
    auto pk = group_plan_kernel<L>;
    int rc = allow_big_smem(pk, "nbg_group(plan): cudaFuncSetAttribute");
    if (rc) return rc;
    const int W = rb_nominal_width<V, CLS>((int)K, g.priv);
    pk<<<(unsigned)g.ntiles, 256, plan_smem, stream>>>(labels, n, (int)K, g.C, g.priv, W, plan);
    rc = check_launch("nbg_group(plan)");
    if (rc) return rc;
    RbParams p;
    p.values = values;
    p.plan = plan;
    p.ws_ch[0] = ws_ch[0], p.ws_ch[1] = ws_ch[1], p.ws_ch[2] = ws_ch[2];
    p.ws_stride = ws_stride;
    p.index_offset = index_offset;
    p.W = W;
    p.locks = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(plan) + ((g.plan_bytes + 15) & ~(size_t)15));
    constexpr bool kPair = CLS == RB_ARGMAX || CLS == RB_ARGMIN || CLS == RB_FIRST || CLS == RB_LAST;
    if (kPair && g.nseg > 1) {
        rc = check_cuda(cudaMemsetAsync(p.locks, 0, g.lock_bytes, stream), "nbg_group(rowbins): lock memset");
        if (rc) return rc;
    }
    p.rows = rows, p.n = n, p.K = (int)K, p.C = g.C;
    p.ntiles = g.ntiles, p.tiles_per_seg = g.tiles_per_seg, p.nseg = g.nseg, p.priv = g.priv;
    const int64_t groups = (rows + kRbRows - 1) / kRbRows;
    if (groups * g.nseg > INT32_MAX) return NBG_OK;
    auto kern = group_rowbins_kernel<V, CLS>;
    rc = allow_big_smem(kern, "nbg_group(rowbins): cudaFuncSetAttribute");
    if (rc) return rc;
    kern<<<(unsigned)(groups * g.nseg), kRbThreads, g.smem, stream>>>(p);
    rc = check_launch("nbg_group(rowbins)");
    if (rc) return rc;
    *handled = true;
    return NBG_OK;
}

}  // namespace nbg
