// nbg_group_rowbins2.cuh -- shared-label grouped reductions, bank-conflict-free edition
// (groupndreduce with axis=int, numbagg/decorators.py:633-645; loop bodies grouped.py:7-270).
//
// Round 1's row-bins kernel (nbg_group_rowbins.cuh) was bound by shared-memory WAVEFRONTS, not by
// HBM: ncu on BASELINE config 2 counted 1.29 G bank conflicts on 1.25 G shared-memory
// instructions -- 9.1 wavefronts per 32 elements against a floor of ~4.5 -- and the kernel time
// equalled wavefronts / 148 SMs to within 1 %.  The conflicts are structural there: TMA places
// every row of a tile on a 16-byte boundary, so element (row, col) can only live in banks
// = col (mod 4) [4-byte data]; four sub-warps reading four unrelated columns collide whenever two
// columns agree mod 4, and their bins collide whenever two labels agree mod 4.
//
// This kernel removes both by construction:
//   * a warp is 4 sub-warps x 8 lanes (lane = row of an 8-row group); sub-warp q only ever
//     processes columns with  col % NCLS == q  (NCLS = 16 / sizeof(V) column classes), and rows
//     are pitched 16 bytes apart in banks, so the 32 lanes of every value load hit 32 distinct
//     banks;
//   * bins are PRIVATE PER COLUMN CLASS: bins[channel][label][class q][row] -- 128 bytes per
//     (channel, label), one bank per lane, any labels -- and the NCLS partial results of a
//     (row, label) are merged when the CTA flushes.  (The summation order therefore differs
//     from the reference's; float sums are compared at rtol 1e-5 / 1e-12 like every other
//     kernel, see DESIGN.md.)
//   * the column plan (built once per call from the labels, shared by all rows) lists, per tile
//     and class, the valid columns stably sorted by label, cut into 64/NCLS label-aligned
//     ranges, each padded to a multiple of 4 entries so that a sub-warp fetches four entries
//     with one 16-byte load, and interleaved so that a group of 4 never holds one label twice:
//     the four read-modify-writes of a group are independent and cost ONE round trip of
//     shared-memory latency (the kernel is latency-bound per warp otherwise).  No atomics
//     anywhere: a label range belongs to one sub-warp per tile and its bins to one lane.
//   * RANGES MOVE ONLY BETWEEN NEIGHBOURS: the ranges of a tile are balanced per tile (every
//     sub-warp gets the same number of entries, so the 4 sub-warps of a warp -- which execute in
//     lock step -- waste no lanes), but boundary j is clamped into a fixed window around its
//     NOMINAL position (cut so that the ranges' shares of all n columns are equal: one global
//     histogram pass over the labels), and the windows of adjacent boundaries do not overlap.
//     A label can therefore only ever belong to range j-1 or j (or j and j+1) -- two sub-warps
//     that live in ADJACENT warps.  A warp starts tile t+1 only after both neighbour warps have
//     finished tile t (two mbarriers per warp, used alternately: arrive = release, try_wait =
//     acquire, hardware sleep instead of polling the shared-memory pipe this kernel is bound
//     by); there is no CTA-wide barrier per tile, and non-adjacent warps drift apart by up to
//     S-1 tiles.
// Row tiles and plan blocks stream through an S-stage TMA ring (cp.async.bulk + mbarrier) filled by
// a dedicated producer warp, which also prefetches the tiles a few stages further ahead into L2
// (cp.async.bulk.prefetch.L2): the bins leave room for only ~90 KB of staging, less than the
// bytes-in-flight HBM latency needs, so DRAM reads are decoupled from the ring.  Consumer warps
// hand stages back through `empty` mbarriers; there is no CTA-wide barrier per tile.
// With few labels (runs of equal labels inside a range) the sub-warp accumulates a run in
// registers and touches the bin once per run (RUNS).
//
// Shared-memory wavefronts per 32 elements: TMA fill 1 + entries 0.25 + values 1 + bins 2 per
// channel = 4.25 (one channel) -- below the 5.6 that HBM allows at 6.45 TB/s and 1.9 GHz, so the
// kernel is HBM-bound for one-channel ops.
#pragma once

#include "nbg_group_rowbins.cuh"

namespace nbg {

constexpr int kRb2MaxWarps = 31;                 // consumer warps per CTA: 16 or 31 (+ the producer warp = 1024 threads)
constexpr int kRb2MaxSub = 4 * kRb2MaxWarps;     // sub-warps = label ranges per tile
constexpr int kRb2Hdr = 136;                     // u32 words per tile header: NCLS * (SUBS + 1) bounds, padded to 16 bytes
constexpr int kRb2Pad = kRb2MaxSub * 3;          // worst-case padding entries per tile
constexpr int kRb2MaxStages = 4;
constexpr int kRb2Header = 640;  // mbarriers: full[4], empty[4], done[31][2]

// ---------------------------------------------------------------------------------- channel ops
// Words have the size of V (float/int32 -> 4 bytes, double/int64 -> 8 bytes); counters are the
// integer of that size (n < 2^31 on this path).
template <typename V>
struct Rb2Word {
    using I = typename RbCounter<V>::type;
};

template <typename V, int CLS>
struct Rb2Op;

#define NBG_RB2_ONE(CLS_, ZERO_, STEP_, MERGE_)                                                        \
    template <typename V>                                                                              \
    struct Rb2Op<V, CLS_> {                                                                            \
        using I = typename Rb2Word<V>::I;                                                              \
        static constexpr int NCH = 1;                                                                  \
        struct W {                                                                                     \
            V a;                                                                                       \
        };                                                                                             \
        __device__ static __forceinline__ W zero() {                                                   \
            W w;                                                                                       \
            ZERO_;                                                                                     \
            return w;                                                                                  \
        }                                                                                              \
        __device__ static __forceinline__ void step(W &w, V v, bool ok) { STEP_; }                     \
        __device__ static __forceinline__ void merge(W &w, const W &o) { MERGE_; }                     \
    }

template <typename V>
__device__ __forceinline__ V rb2_from_int(typename Rb2Word<V>::I i) {
    V v;
    memcpy(&v, &i, sizeof(V));
    return v;
}
template <typename V>
__device__ __forceinline__ typename Rb2Word<V>::I rb2_to_int(V v) {
    typename Rb2Word<V>::I i;
    memcpy(&i, &v, sizeof(V));
    return i;
}

NBG_RB2_ONE(RB_SUM, w.a = (V)0, w.a = v_add(w.a, ok ? v : (V)0), w.a = v_add(w.a, o.a));
NBG_RB2_ONE(RB_SUMSQ, w.a = (V)0, w.a = v_add(w.a, v_sq(ok ? v : (V)0)), w.a = v_add(w.a, o.a));
NBG_RB2_ONE(RB_PROD, w.a = (V)1, w.a = v_mul(w.a, ok ? v : (V)1), w.a = v_mul(w.a, o.a));
// integer-valued words live in the V-sized slot as raw bits
NBG_RB2_ONE(RB_COUNT, w.a = rb2_from_int<V>(0), w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) + (ok ? 1 : 0)),
            w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) + rb2_to_int<V>(o.a)));
NBG_RB2_ONE(RB_ANY, w.a = rb2_from_int<V>(0), w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) | ((ok && v != (V)0) ? 1 : 0)),
            w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) | rb2_to_int<V>(o.a)));
NBG_RB2_ONE(RB_ALL, w.a = rb2_from_int<V>(1), w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) & ((ok && v == (V)0) ? 0 : 1)),
            w.a = rb2_from_int<V>(rb2_to_int<V>(w.a) & rb2_to_int<V>(o.a)));
#undef NBG_RB2_ONE

template <typename V>
struct Rb2Op<V, RB_MEAN> {  // grouped.py:7-27
    using I = typename Rb2Word<V>::I;
    static constexpr int NCH = 2;
    struct W {
        V a;
        V b;  // count bits
    };
    __device__ static __forceinline__ W zero() { return W{(V)0, rb2_from_int<V>(0)}; }
    __device__ static __forceinline__ void step(W &w, V v, bool ok) {
        w.a = v_add(w.a, ok ? v : (V)0);
        w.b = rb2_from_int<V>(rb2_to_int<V>(w.b) + (ok ? 1 : 0));
    }
    __device__ static __forceinline__ void merge(W &w, const W &o) {
        w.a = v_add(w.a, o.a);
        w.b = rb2_from_int<V>(rb2_to_int<V>(w.b) + rb2_to_int<V>(o.b));
    }
};
template <typename V>
struct Rb2Op<V, RB_VAR> {  // grouped.py:148-208
    using I = typename Rb2Word<V>::I;
    static constexpr int NCH = 3;
    struct W {
        V a, b;
        V c;  // count bits
    };
    __device__ static __forceinline__ W zero() { return W{(V)0, (V)0, rb2_from_int<V>(0)}; }
    __device__ static __forceinline__ void step(W &w, V v, bool ok) {
        const V m = ok ? v : (V)0;
        w.a = v_add(w.a, m);
        w.b = v_add(w.b, v_sq(m));
        w.c = rb2_from_int<V>(rb2_to_int<V>(w.c) + (ok ? 1 : 0));
    }
    __device__ static __forceinline__ void merge(W &w, const W &o) {
        w.a = v_add(w.a, o.a);
        w.b = v_add(w.b, o.b);
        w.c = rb2_from_int<V>(rb2_to_int<V>(w.c) + rb2_to_int<V>(o.c));
    }
};

// merged words -> the round-1 bin type, so that its workspace flush (RbFlush) is reused as is
template <typename V, int CLS>
__device__ __forceinline__ RbBin<V, CLS> rb2_to_bin(const typename Rb2Op<V, CLS>::W &w) {
    RbBin<V, CLS> b;
    if constexpr (CLS == RB_SUM || CLS == RB_SUMSQ) {
        b.s = w.a;
    } else if constexpr (CLS == RB_PROD) {
        b.pr = w.a;
    } else if constexpr (CLS == RB_COUNT) {
        b.c = rb2_to_int<V>(w.a);
    } else if constexpr (CLS == RB_ANY || CLS == RB_ALL) {
        b.flag = rb2_to_int<V>(w.a);
    } else if constexpr (CLS == RB_MEAN) {
        b.s = w.a;
        b.c = rb2_to_int<V>(w.b);
    } else {
        b.s = w.a;
        b.ss = w.b;
        b.c = rb2_to_int<V>(w.c);
        b.pad = 0;
    }
    return b;
}

// ----------------------------------------------------------------------------------- plan
// Per tile: header hdr[q * (SUBS + 1) + j] = first entry of range j of class q (the next range
// starts where this one ends; every range is a multiple of 4 entries long), then the entries
// (byte offset of the column inside a tile row << 19 | label * 128) -- both fields pre-scaled (13 + 19
// bits: tile rows up to 8 KB, up to 4094 labels) so that the kernel forms each shared-memory address
// with one or two integer instructions.  Padding entries
// point at the dummy slot `nslots` and at column q (same class, always inside the tile).  A CTA of a cluster of `nc` owns the labels with
// label % nc == rank and addresses them as slot = label / nc; its plan lists only those.
// Global label histogram per column class (hist[class * K + label], zeroed by the caller).
template <typename L, int NCLS>
__global__ void __launch_bounds__(256) group_hist2_kernel(const L *__restrict__ labels, int64_t n, int K,
                                                          int *__restrict__ hist) {
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = i0 + k;
        if (i >= n) break;
        const long long lab = (long long)labels[i];
        if (lab >= 0 && lab < K) atomicAdd(&hist[(int)(i % NCLS) * K + (int)lab], 1);
    }
}
// cuts[q * (G + 1) + g] = first label of group g's interval in class q (g = 0..G; [G] = K): the
// smallest label whose exclusive prefix count reaches g/G of the class total.
__global__ void __launch_bounds__(256) group_cuts2_kernel(const int *__restrict__ hist, int K, int ncls, int G,
                                                          int *__restrict__ cuts) {
    extern __shared__ int pre[];  // [K + 1] exclusive prefix of one class
    __shared__ int wtot[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int q = 0; q < ncls; q++) {
        const int per = (K + 255) / 256;
        const int beg = min(tid * per, K), end = min(beg + per, K);
        int local = 0;
        for (int k = beg; k < end; k++) local += hist[q * K + k];
        int inc = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        int base = 0;
        for (int w2 = 0; w2 < wid; w2++) base += wtot[w2];
        int run = base + inc - local;
        for (int k = beg; k < end; k++) {
            pre[k] = run;
            run += hist[q * K + k];
        }
        if (tid == 255) pre[K] = run;
        __syncthreads();
        if (tid <= G) {
            int cut = K;
            if (tid < G) {
                const long long target = ((long long)tid * pre[K]) / G;
                int lo = 0, hi = K;  // first label with pre[label] >= target
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (pre[mid] >= target) hi = mid;
                    else lo = mid + 1;
                }
                cut = tid == 0 ? 0 : lo;
            }
            cuts[q * (G + 1) + tid] = cut;
        }
        __syncthreads();
    }
}

template <typename L, int NCLS>
__global__ void __launch_bounds__(256) group_plan2_kernel(const L *__restrict__ labels, int64_t n, int K, int C,
                                                          int ent_cap, int nc, int vsize, int runs, int nsub, const int *__restrict__ cuts,
                                                          uint32_t *__restrict__ plan) {
    const int SUBS = nsub / NCLS;  // ranges per class
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NK = NCLS * K;
    int *offs = reinterpret_cast<int *>(smem_raw);               // [NK + 1] histogram -> exclusive offsets
    int *cursor = offs + (NK + 1);                               // [NK]
    uint32_t *skey = reinterpret_cast<uint32_t *>(cursor + NK);  // [C] class * K + label, or ~0
    uint32_t *sent = skey + C;                                   // [C] entries in (class, label, column) order
    __shared__ int warp_tot[8];
    __shared__ int ub[kRb2MaxSub + 1];   // unpadded range bounds over `sent`, flattened (class, range)
    __shared__ int pst[kRb2MaxSub + 1];  // padded start of every range
    __shared__ int rseq[kRb2MaxSub];     // range keeps sorted order (sequential read-modify-writes)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x / nc, rank = blockIdx.x % nc;
    const int64_t c0 = (int64_t)tile * C;
    uint32_t *out = plan + (size_t)blockIdx.x * (kRb2Hdr + ent_cap);
    const int nslots = (K + nc - 1) / nc;

    for (int k = tid; k <= NK; k += 256) offs[k] = 0;
    __syncthreads();
    for (int j = tid; j < C; j += 256) {
        const int64_t col = c0 + j;
        long long lab = -1;
        if (col < n) lab = (long long)labels[col];
        const bool valid = lab >= 0 && lab < K && (nc == 1 || (int)(lab % nc) == rank);
        const uint32_t key = (uint32_t)((j % NCLS) * K + (int)lab);
        skey[j] = valid ? key : 0xffffffffu;
        if (valid) atomicAdd(&offs[key], 1);
    }
    __syncthreads();
    {  // exclusive scan of offs[0..NK)
        const int per = (NK + 255) / 256;
        const int beg = min(tid * per, NK), end = min(beg + per, NK);
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
            offs[NK] = tot;
        }
    }
    __syncthreads();
    // stable placement by one warp: columns in ascending order, 32 at a time
    if (wid == 0) {
        for (int j0 = 0; j0 < C; j0 += 32) {
            const int j = j0 + lane;
            const uint32_t key = j < C ? skey[j] : 0xffffffffu;
            const bool valid = key != 0xffffffffu;
            const unsigned m = __match_any_sync(0xffffffffu, key);
            if (valid) {
                const int rk = __popc(m & ((1u << lane) - 1u));
                const int lab = (int)(key % (uint32_t)K);
                sent[cursor[key] + rk] = ((uint32_t)(j * vsize) << 19) | ((uint32_t)(lab / nc) << 7);
            }
            __syncwarp();
            if (valid && (m & ((1u << lane) - 1u)) == 0) cursor[key] += __popc(m);
            __syncwarp();
        }
    }
    __syncthreads();
    // ranges (unpadded positions), balanced per tile; the first label of range j is clamped into
    // [M(j-1), M(j)], M(j) = midpoint of the nominal cuts j and j+1 (cuts[q][0] = 0, cuts[q][SUBS] = K)
    if (tid <= nsub) {
        int b;
        if (tid == nsub) {
            b = offs[NK];
        } else {
            const int q = tid / SUBS, j = tid % SUBS;
            const int cs = offs[q * K], ce = offs[(q + 1) * K];
            if (j == 0) {
                b = cs;
            } else {
                const int *cq = cuts + q * (SUBS + 1);
                const int m_lo = (cq[j - 1] + cq[j]) >> 1, m_hi = (cq[j] + cq[j + 1]) >> 1;
                int lab = cq[j];
                if (ce > cs) {
                    const int e = cs + (int)(((long long)j * (ce - cs)) / SUBS);
                    lab = (int)((sent[e] & 0x7ffffu) >> 7) * nc + rank;  // label of the entry at the even split
                }
                lab = max(m_lo, min(lab, m_hi));
                b = offs[q * K + lab];  // first entry of that label in this tile (lab == K: end of class)
            }
        }
        ub[tid] = b;
    }
    __syncthreads();
    if (tid == 0) {
        int p = 0;
        for (int g = 0; g < nsub; g++) {
            pst[g] = p;
            p += (ub[g + 1] - ub[g] + 3) & ~3;
        }
        pst[nsub] = p;
    }
    __syncthreads();
    // Layout inside a range of c entries, G = ceil(c / 4) groups of 4: sorted position p goes to slot
    // 4 * (p % G) + p / G, i.e. a group holds the sorted positions {g, G+g, 2G+g, 3G+g}.  Entries of one
    // label are contiguous in sorted order, so as long as no label has more than G entries in the range a
    // group never holds a label twice and its four read-modify-writes are INDEPENDENT (the kernel issues
    // them together).  Otherwise -- or with `runs` (few labels: register accumulation needs sorted order)
    // -- the range keeps the sorted order and is flagged sequential (bit 31 of its header word).
    if (tid < nsub) {
        const int q = tid / SUBS;
        const int cnt = ub[tid + 1] - ub[tid];
        const int G = (cnt + 3) >> 2;
        int seq = runs;
        if (!seq && cnt > 0) {
            // longest label run of this range
            const int l0 = (int)((sent[ub[tid]] & 0x7ffffu) >> 7) * nc + rank, l1 = (int)((sent[ub[tid + 1] - 1] & 0x7ffffu) >> 7) * nc + rank;
            int longest = 0;
            for (int lab = l0; lab <= l1; lab++) {
                const int a0 = max(offs[q * K + lab], ub[tid]), a1 = min(offs[q * K + lab + 1], ub[tid + 1]);
                longest = max(longest, a1 - a0);
            }
            seq = longest > G;
        }
        rseq[tid] = seq;
    }
    __syncthreads();
    // header: hdr[q * (SUBS + 1) + j], j = 0..SUBS  (entry SUBS of class q = start of class q + 1)
    if (tid < NCLS * (SUBS + 1)) {
        const int q = tid / (SUBS + 1), j = tid % (SUBS + 1);
        const int r = q * SUBS + j;
        out[tid] = (uint32_t)pst[r] | ((j < SUBS && rseq[r]) ? 0x80000000u : 0u);
    }
    for (int t2 = NCLS * (SUBS + 1) + tid; t2 < kRb2Hdr; t2 += 256) out[t2] = 0;
    // dummies first (a range's padding slots are scattered in the interleaved layout), then the entries
    if (tid < nsub) {
        const int cnt = ub[tid + 1] - ub[tid];
        const int q = tid / SUBS;
        const uint32_t dummy = ((uint32_t)(q * vsize) << 19) | ((uint32_t)nslots << 7);
        if (cnt & 3) {
            const int m = (cnt + 3) & ~3;
            for (int i = max(0, m - 16); i < m; i++) out[kRb2Hdr + pst[tid] + i] = dummy;  // padding lives in the last 4 groups
        }
    }
    __syncthreads();
    const int total = offs[NK];
    for (int pos = tid; pos < total; pos += 256) {
        int lo = 0, hi = nsub - 1;  // last range g with ub[g] <= pos
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ub[mid] <= pos) lo = mid;
            else hi = mid - 1;
        }
        const int p = pos - ub[lo];
        int slot = p;
        if (!rseq[lo]) {
            const int G = (ub[lo + 1] - ub[lo] + 3) >> 2;
            slot = 4 * (p % G) + p / G;
        }
        out[kRb2Hdr + pst[lo] + slot] = sent[pos];
    }
}

// ----------------------------------------------------------------------------------- kernel
struct Rb2Params {
    const void *values;
    const uint32_t *plan;
    void *ws_ch[3];
    int64_t ws_stride;
    int64_t rows, n;
    int K, C, S, ent_cap;
    int ntiles, tiles_per_seg, nseg;
    int PD;  // tiles beyond the ring that the producer prefetches into L2 (0: off)
    int nc;      // label split: nc CTAs share a row group, CTA `rank` owns the labels with label % nc == rank
    int nslots;  // ceil(K / nc) bins per CTA
    int64_t index_offset;
};

template <typename V>
__host__ __device__ inline size_t rb2_stage_bytes(int C, int ent_cap) {
    return ((size_t)(kRb2Hdr + ent_cap) * 4 + (size_t)kRbRows * rb_row_stride<V>(C) * sizeof(V) + 15) & ~(size_t)15;
}
template <typename V, int CLS>
__host__ __device__ inline size_t rb2_bins_bytes(int nslots) {
    return (size_t)Rb2Op<V, CLS>::NCH * (size_t)(nslots + 1) * 128;
}
template <typename V, int CLS>
__host__ __device__ inline size_t rb2_smem_bytes(int nslots, int C, int ent_cap, int S) {
    return kRb2Header + rb2_bins_bytes<V, CLS>(nslots) + (size_t)S * rb2_stage_bytes<V>(C, ent_cap);
}

template <typename V, int CLS, bool RUNS, int NW>
__global__ void __launch_bounds__(NW * 32 + 32, 1) group_rowbins2_kernel(Rb2Params p) {
    constexpr int kRb2Consumers = NW * 32;            // consumer threads = NW * 4 sub-warps of 8 rows
    constexpr int kRb2Threads = kRb2Consumers + 32;   // + one producer warp
    using Op = Rb2Op<V, CLS>;
    using W = typename Op::W;
    using Acc = typename std::conditional<std::is_floating_point<V>::value, double, long long>::type;
    constexpr int NCH = Op::NCH;
    constexpr int NCLS = 16 / (int)sizeof(V);
    constexpr int SUBS = NW * 4 / NCLS;
    static_assert(sizeof(W) == NCH * sizeof(V), "channel words are V-sized");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);       // [S]
    uint64_t *empty = full + kRb2MaxStages;                        // [S]
    uint64_t *done = empty + kRb2MaxStages;                        // [16][2]: warp w finished a tile of parity b
    unsigned char *bins = smem_raw + kRb2Header;
    const int C = p.C, K = p.K, S = p.S;
    const int nc = p.nc, rank = (int)(blockIdx.x % (unsigned)nc);
    const int nslots = p.nslots;  // this CTA's labels: rank, rank + nc, ... (slot = label / nc)
    const size_t bins_bytes = rb2_bins_bytes<V, CLS>(nslots);
    const size_t ch_bytes = (size_t)(nslots + 1) * 128;
    unsigned char *stage0 = bins + bins_bytes;
    const size_t stage_bytes = rb2_stage_bytes<V>(C, p.ent_cap);
    const int stride = rb_row_stride<V>(C);

    const int tid = threadIdx.x;
    const int64_t gs = blockIdx.x / (unsigned)nc;  // the nc CTAs of a row group are adjacent: they stream the same tiles
    const int64_t g = gs / p.nseg;
    const int seg = (int)(gs % p.nseg);
    const int64_t r0 = g * kRbRows;
    const int nrows = (int)min((int64_t)kRbRows, p.rows - r0);
    const int t_beg = seg * p.tiles_per_seg;
    const int t_end = min(t_beg + p.tiles_per_seg, p.ntiles);
    const int ntl = t_end - t_beg;
    const V *vbase = reinterpret_cast<const V *>(p.values) + r0 * p.n;

    // bins start at the identity of every channel (word images are channel-uniform)
    {
        const W z = Op::zero();
        const V *zw = reinterpret_cast<const V *>(&z);
        constexpr int WPS = 128 / (int)sizeof(V);  // words per (channel, slot)
        V *bw = reinterpret_cast<V *>(bins);
        for (int ch = 0; ch < NCH; ch++)
            for (int i = tid; i < (nslots + 1) * WPS; i += kRb2Threads) bw[(size_t)ch * (nslots + 1) * WPS + i] = zw[ch];
    }
    if (tid == 0) {
        for (int w2 = 0; w2 < 2 * (kRb2Consumers / 32); w2++) mbar_init(&done[w2], 1);
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kRb2Consumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kRb2Consumers) {
        // ---------------------------------------------------------------- producer warp
        if (tid == kRb2Consumers) {
            const uint32_t plan_bytes = (uint32_t)(kRb2Hdr + p.ent_cap) * 4u;
            int st = 0;
            uint32_t ph = 1;  // fresh barrier: the first pass over the ring does not wait
            for (int it = 0; it < ntl; it++, st = (st + 1 == S ? 0 : st + 1), ph ^= (st == 0)) {
                mbar_wait(&empty[st], ph);
                const int t = t_beg + it;
                unsigned char *sb = stage0 + (size_t)st * stage_bytes;
                const int64_t c0 = (int64_t)t * C;
                const int cols = (int)min((int64_t)C, p.n - c0);
                const uint32_t row_bytes = (uint32_t)cols * (uint32_t)sizeof(V);
                mbar_arrive_expect_tx(&full[st], plan_bytes + (uint32_t)nrows * row_bytes);
                bulk_g2s(sb, p.plan + ((size_t)t * nc + rank) * (kRb2Hdr + p.ent_cap), plan_bytes, &full[st]);
                V *tile = reinterpret_cast<V *>(sb + (size_t)(kRb2Hdr + p.ent_cap) * 4);
                for (int rr = 0; rr < nrows; rr++)
                    bulk_g2s(tile + (size_t)rr * stride, vbase + (int64_t)rr * p.n + c0, row_bytes, &full[st]);
                // L2 prefetch of the tile PD steps ahead (first step: everything up to it)
                if (p.PD > 0) {
                    for (int t2 = (it == 0 ? t + 1 : t + p.PD); t2 <= t + p.PD && t2 < t_end; t2++) {
                        const int64_t d0 = (int64_t)t2 * C;
                        const uint32_t db = (uint32_t)min((int64_t)C, p.n - d0) * (uint32_t)sizeof(V);
                        for (int rr = 0; rr < nrows; rr++) bulk_prefetch_l2(vbase + (int64_t)rr * p.n + d0, db);
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- consumer warps
        const int r = tid & (kRbRows - 1);
        const int sub = tid >> 3;
        const int q = sub % NCLS;   // column class of this sub-warp
        const int j = sub / NCLS;   // its label range within the class
        const uint32_t bins_s = smem_u32(bins) + (uint32_t)(q * kRbRows + r) * (uint32_t)sizeof(V);
        const bool active = r < nrows;
        const int wrp = tid >> 5;
        int st = 0;
        uint32_t ph = 0;
        for (int it = 0; it < ntl; it++, st = (st + 1 == S ? 0 : st + 1), ph ^= (st == 0)) {
            if (it > 0) {
                // labels move only between ADJACENT ranges from one tile to the next: both neighbour
                // warps must have finished the previous tile before this one touches its bins
                // (a neighbour is never more than one tile ahead of this wait, so the two barriers of a
                // warp cannot alias phases)
                const int tp = it - 1;
                const uint32_t par = (uint32_t)((tp >> 1) & 1);
                if (wrp > 0) mbar_wait(&done[2 * (wrp - 1) + (tp & 1)], par);
                if (wrp < kRb2Consumers / 32 - 1) mbar_wait(&done[2 * (wrp + 1) + (tp & 1)], par);
            }
            mbar_wait(&full[st], ph);
            const uint32_t sb = smem_u32(stage0 + (size_t)st * stage_bytes);
            const uint32_t *hdr = reinterpret_cast<const uint32_t *>(__cvta_shared_to_generic(sb));
            const uint32_t ent_s = sb + kRb2Hdr * 4;
            const uint32_t tile_s = sb + (uint32_t)(kRb2Hdr + p.ent_cap) * 4 + (uint32_t)r * (uint32_t)stride * (uint32_t)sizeof(V);
            const uint32_t h0 = hdr[q * (SUBS + 1) + j];
            const int b0 = (int)(h0 & 0x7fffffffu), b1 = (int)(hdr[q * (SUBS + 1) + j + 1] & 0x7fffffffu);
            const bool sequential = (h0 >> 31) != 0;  // a label occurs twice inside some group of 4 (or RUNS)
            auto val_at = [&](uint32_t e) -> V {
                return *reinterpret_cast<const V *>(__cvta_shared_to_generic(tile_s + (e >> 19)));
            };
            auto bin_ptr = [&](uint32_t slot, int ch) -> V * {
                return reinterpret_cast<V *>(__cvta_shared_to_generic(bins_s + (uint32_t)ch * (uint32_t)ch_bytes + slot));  // `slot` is pre-scaled: label * 128
            };
            auto rmw = [&](uint32_t slot, V v) {
                W w;
                V *ww = reinterpret_cast<V *>(&w);
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) ww[ch] = *bin_ptr(slot, ch);
                Op::step(w, v, !is_nan(v));
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) *bin_ptr(slot, ch) = ww[ch];
            };
            if (active) {
                if constexpr (!RUNS) {
                    if (!sequential) {
                        // the plan guarantees four DIFFERENT labels per group: load the four bins together,
                        // update, store -- one round trip of shared-memory latency per group instead of four
                        // software pipeline: the NEXT group's entries and values (read-only data, cannot
                        // alias the bins) are fetched before this group's read-modify-write
                        auto ld_ent = [&](int i) -> uint4 {
                            return *reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(ent_s + (uint32_t)i * 4u));
                        };
                        uint4 e = make_uint4(0, 0, 0, 0);
                        V v0 = (V)0, v1 = (V)0, v2 = (V)0, v3 = (V)0;
                        if (b0 < b1) {
                            e = ld_ent(b0);
                            v0 = val_at(e.x), v1 = val_at(e.y), v2 = val_at(e.z), v3 = val_at(e.w);
                        }
                        for (int i = b0; i < b1; i += 4) {
                            uint4 en = e;
                            V n0 = v0, n1 = v1, n2 = v2, n3 = v3;
                            if (i + 4 < b1) {
                                en = ld_ent(i + 4);
                                n0 = val_at(en.x), n1 = val_at(en.y), n2 = val_at(en.z), n3 = val_at(en.w);
                            }
                            const uint32_t s0 = e.x & 0x7ffffu, s1 = e.y & 0x7ffffu, s2 = e.z & 0x7ffffu, s3 = e.w & 0x7ffffu;
                            W w0, w1, w2, w3;
                            V *p0 = reinterpret_cast<V *>(&w0), *p1 = reinterpret_cast<V *>(&w1);
                            V *p2 = reinterpret_cast<V *>(&w2), *p3 = reinterpret_cast<V *>(&w3);
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) {
                                p0[ch] = *bin_ptr(s0, ch);
                                p1[ch] = *bin_ptr(s1, ch);
                                p2[ch] = *bin_ptr(s2, ch);
                                p3[ch] = *bin_ptr(s3, ch);
                            }
                            Op::step(w0, v0, !is_nan(v0));
                            Op::step(w1, v1, !is_nan(v1));
                            Op::step(w2, v2, !is_nan(v2));
                            Op::step(w3, v3, !is_nan(v3));
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) {
                                *bin_ptr(s0, ch) = p0[ch];
                                *bin_ptr(s1, ch) = p1[ch];
                                *bin_ptr(s2, ch) = p2[ch];
                                *bin_ptr(s3, ch) = p3[ch];
                            }
                            e = en;
                            v0 = n0, v1 = n1, v2 = n2, v3 = n3;
                        }
                    } else {
                        for (int i = b0; i < b1; i += 4) {
                            const uint4 e = *reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(ent_s + (uint32_t)i * 4u));
                            const V v0 = val_at(e.x), v1 = val_at(e.y), v2 = val_at(e.z), v3 = val_at(e.w);
                            // read-modify-write in entry order: consecutive entries may share a label
                            rmw(e.x & 0x7ffffu, v0);
                            rmw(e.y & 0x7ffffu, v1);
                            rmw(e.z & 0x7ffffu, v2);
                            rmw(e.w & 0x7ffffu, v3);
                        }
                    }
                } else {
                    uint32_t cur = 0xffffffffu;
                    W acc = Op::zero();
                    auto flush_run = [&]() {
                        W w;
                        V *ww = reinterpret_cast<V *>(&w);
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++) ww[ch] = *bin_ptr(cur, ch);
                        Op::merge(w, acc);
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++) *bin_ptr(cur, ch) = ww[ch];
                    };
                    auto one = [&](uint32_t e, V v) {
                        const uint32_t slot = e & 0x7ffffu;
                        if (slot != cur) {
                            if (cur != 0xffffffffu) flush_run();
                            cur = slot;
                            acc = Op::zero();
                        }
                        Op::step(acc, v, !is_nan(v));
                    };
                    for (int i = b0; i < b1; i += 4) {
                        const uint4 e = *reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(ent_s + (uint32_t)i * 4u));
                        const V v0 = val_at(e.x), v1 = val_at(e.y), v2 = val_at(e.z), v3 = val_at(e.w);
                        one(e.x, v0);
                        one(e.y, v1);
                        one(e.z, v2);
                        one(e.w, v3);
                    }
                    if (cur != 0xffffffffu) flush_run();
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) {
                mbar_arrive(&done[2 * wrp + (it & 1)]);  // release: this warp's bin updates of tile `it`
                mbar_arrive(&empty[st]);                 // and it is done with stage st
            }
        }
    }
    __syncthreads();

    // ---- flush: merge the NCLS class-private partials of every (row, label), then into the workspace
    Acc *c0p = reinterpret_cast<Acc *>(p.ws_ch[0]);
    Acc *c1p = reinterpret_cast<Acc *>(p.ws_ch[1]);
    long long *c2p = reinterpret_cast<long long *>(p.ws_ch[2]);
    const bool atomic = p.nseg > 1;
    constexpr int WPS = 128 / (int)sizeof(V);
    const V *bw = reinterpret_cast<const V *>(bins);
    for (int idx = tid; idx < nslots * kRbRows; idx += kRb2Threads) {
        const int rr = idx & (kRbRows - 1), slot = idx >> 3;  // row fastest: 8 lanes read 8 neighbouring banks
        const int k = slot * nc + rank;                        // the label behind this CTA's slot
        if (rr >= nrows || k >= K) continue;
        W tot;
        V *tw = reinterpret_cast<V *>(&tot);
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) tw[ch] = bw[((size_t)ch * (nslots + 1) + slot) * WPS + rr];
#pragma unroll
        for (int qq = 1; qq < NCLS; qq++) {
            W o;
            V *ow = reinterpret_cast<V *>(&o);
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) ow[ch] = bw[((size_t)ch * (nslots + 1) + slot) * WPS + qq * kRbRows + rr];
            Op::merge(tot, o);
        }
        const RbBin<V, CLS> b = rb2_to_bin<V, CLS>(tot);
        const size_t o = ((size_t)(r0 + rr) * K + k) * (size_t)p.ws_stride;  // record offset (words)
        if constexpr (CLS <= RB_VAR) {
            RbFlush<V, CLS>::flush(b, c0p + o, c1p + o, c2p + o, atomic);
        } else {
            RbFlush<V, CLS>::flush(b, c0p + o, c1p + o, p.index_offset, atomic);
        }
    }
}

// ------------------------------------------------------------------------------------ host
struct Rb2Geometry {
    bool ok;
    int C, S, ent_cap, ntiles, nseg, tiles_per_seg, runs, PD, NW, nc, nslots;
    size_t smem, plan_bytes, plan_smem, aux_bytes;
};

template <typename V, int CLS>
static Rb2Geometry rb2_geometry(int64_t rows, int64_t n, int64_t K) {
    Rb2Geometry g = {};
    constexpr int PER16 = 16 / (int)sizeof(V);
    constexpr int NCLS = PER16;
    if (K <= 0 || K > 4094 || n <= 0 || n >= ((int64_t)1 << 31) || (n % PER16) != 0 || rows < 1) return g;
    if (getenv("NBG_RB2_OFF")) return g;
    // Label split (NBG_RB2_NC, experiment): the class-private bins cost 128 bytes per (channel, label), so
    // mean / var at 1000 float32 labels (256 / 384 KB) do not fit one CTA.  nc CTAs can share a row group
    // -- adjacent in the grid, streaming the same tiles, CTA `rank` owning the labels with
    // label % nc == rank -- but every one of them still stages whole tiles and pays the per-tile costs
    // for 1/nc of the entries.  Measured on config 2: nansum 8.4 ms (nc = 1) / 10.2 (2) / 18.5 (4);
    // nanmean 15.1 (nc = 2) against 13.7 ms for the round-1 kernel; nanstd 23.9 (nc = 3) against 20.9.
    // Off by default: ops whose bins do not fit keep the round-1 kernel.
    int nc = 1;
    if (const char *e = getenv("NBG_RB2_NC")) nc = atoi(e) > 0 ? atoi(e) : 1;
    if (nc > K) nc = (int)K;
    const int nslots = (int)((K + nc - 1) / nc);
    const size_t bins = rb2_bins_bytes<V, CLS>(nslots);
    if (bins + kRb2Header > kMaxSmemOptIn) return g;
    const size_t avail = kMaxSmemOptIn - kRb2Header - bins;
    g.nc = nc, g.nslots = nslots;
    int S = 2, C = 0;  // two wide stages beat three narrower ones (config 2: 8.6 vs 9.2 ms): per-tile costs dominate
    if (const char *e = getenv("NBG_RB2_S")) S = atoi(e);
    if (S < 2) S = 2;
    if (S > kRb2MaxStages) S = kRb2MaxStages;
    auto fit = [&](int s) {
        int c = 8192 / (int)sizeof(V) - 64;  // the entry's 13-bit byte offset
        while (c >= 64 && (size_t)s * rb2_stage_bytes<V>(c, c + kRb2Pad) > avail) c -= 64;
        return c >= 64 ? c : 0;
    };
    C = fit(S);
    if (C < 256 && S > 2) {
        S = 2;
        C = fit(S);
    }
    if (const char *e = getenv("NBG_RB2_C")) C = atoi(e);
    if (C < 128 || (C % PER16) != 0) return g;
    // tiles never wider than the row (short rows: one tile)
    const int64_t n_up = (n + 63) / 64 * 64;
    if (C > n_up) C = (int)n_up;
    g.C = C, g.S = S, g.ent_cap = C + kRb2Pad;
    g.smem = rb2_smem_bytes<V, CLS>(nslots, C, g.ent_cap, S);
    if (g.smem > kMaxSmemOptIn) return g;
    g.plan_smem = (size_t)(2 * NCLS * K + 1) * 4 + (size_t)2 * C * 4 + 16;
    if (g.plan_smem > kMaxSmem) return g;
    g.ntiles = (int)((n + C - 1) / C);
    const int64_t groups = (rows + kRbRows - 1) / kRbRows;
    // one CTA per SM: whole rows per CTA once there are >= 4 waves of row groups, otherwise column
    // segments (~8 waves) merged into the workspace with atomics
    const int64_t slots = kNumSMs;
    int64_t nseg = groups * nc >= 4 * slots ? 1 : (8 * slots + groups * nc - 1) / (groups * nc);
    if (const char *e = getenv("NBG_RB2_NSEG")) nseg = atoi(e);
    if (nseg > g.ntiles) nseg = g.ntiles;
    if (nseg < 1) nseg = 1;
    g.tiles_per_seg = (int)((g.ntiles + nseg - 1) / nseg);
    g.nseg = (g.ntiles + g.tiles_per_seg - 1) / g.tiles_per_seg;
    g.plan_bytes = (size_t)g.ntiles * nc * (kRb2Hdr + g.ent_cap) * 4 + 256;
    // average run of equal labels inside a class range >= 2: accumulate runs in registers
    g.runs = ((int64_t)C / NCLS >= 2 * K) ? 1 : 0;
    if (const char *e = getenv("NBG_RB2_RUNS")) g.runs = atoi(e);
    g.PD = 0;  // measured: no gain on config 2 (the ring already covers the latency at this rate)
    if (const char *e = getenv("NBG_RB2_PD")) g.PD = atoi(e);
    // global histogram per class + group cuts (ints)
    g.aux_bytes = ((size_t)NCLS * K + (size_t)NCLS * 65 + 64) * 4;
    // consumer warps: 16.  31 (+ producer = 1024 threads) was measured SLOWER on config 2 (10.2 vs 8.6 ms):
    // half-sized ranges double the padding and the per-tile overhead (+28 % instructions); 8 / 12 / 24 warps
    // measured 8.69 / 8.35 / 9.25 ms against 8.43 (profiles/experiments/r02_exp29_cfg2.jsonl): a plateau
    g.NW = 16;
    if (const char *e = getenv("NBG_RB2_NW")) g.NW = atoi(e) == 31 ? 31 : 16;
    g.ok = true;
    return g;
}

inline size_t rb2_scratch_bytes(int64_t n) {
    // narrowest tile is 128 columns: header + padding per tile, 4 bytes per column -- times the label
    // split (every rank's plan block has room for a whole tile; up to 4 ranks fit, else the split is refused)
    return (size_t)(n + 2048) * 16 + (size_t)(n / 128 + 4) * (kRb2Hdr + kRb2Pad) * 4 + 4096 + ((size_t)4 * 65536 + 256) * 4;
}

template <typename V, typename L, int CLS>
static int rb2_launch(const V *values, const L *labels, void *ws_ch[3], int64_t ws_stride, void *scratch, size_t scratch_bytes,
                      int64_t rows, int64_t n, int64_t K, int64_t index_offset, cudaStream_t stream, bool *handled) {
    *handled = false;
    constexpr int NCLS = 16 / (int)sizeof(V);
    const Rb2Geometry g = rb2_geometry<V, CLS>(rows, n, K);
    if (!g.ok || scratch == nullptr || scratch_bytes < g.plan_bytes + g.aux_bytes + 512) return NBG_OK;
    if (((uintptr_t)values & 15) != 0) return NBG_OK;
    uint32_t *plan = reinterpret_cast<uint32_t *>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    int *hist = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(plan) + ((g.plan_bytes + 255) & ~(size_t)255));
    int *cuts = hist + (size_t)NCLS * K;
    const int SUBS = g.NW * 4 / NCLS;  // ranges per class
    // labels only: global per-class histogram -> group cuts -> per-tile plan
    int rc = check_cuda(cudaMemsetAsync(hist, 0, (size_t)NCLS * K * sizeof(int), stream), "nbg_group(hist2): memset");
    if (rc) return rc;
    group_hist2_kernel<L, NCLS><<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(labels, n, (int)K, hist);
    rc = check_launch("nbg_group(hist2)");
    if (rc) return rc;
    group_cuts2_kernel<<<1, 256, (size_t)(K + 1) * sizeof(int), stream>>>(hist, (int)K, NCLS, SUBS, cuts);
    rc = check_launch("nbg_group(cuts2)");
    if (rc) return rc;
    auto pk = group_plan2_kernel<L, NCLS>;
    rc = allow_big_smem(pk, "nbg_group(plan2): cudaFuncSetAttribute");
    if (rc) return rc;
    pk<<<(unsigned)(g.ntiles * g.nc), 256, g.plan_smem, stream>>>(labels, n, (int)K, g.C, g.ent_cap, g.nc, (int)sizeof(V), g.runs, g.NW * 4, cuts, plan);
    rc = check_launch("nbg_group(plan2)");
    if (rc) return rc;
    Rb2Params p;
    p.values = values;
    p.plan = plan;
    p.ws_ch[0] = ws_ch[0], p.ws_ch[1] = ws_ch[1], p.ws_ch[2] = ws_ch[2];
    p.ws_stride = ws_stride;
    p.index_offset = index_offset;
    p.rows = rows, p.n = n, p.K = (int)K, p.C = g.C, p.S = g.S, p.ent_cap = g.ent_cap;
    p.ntiles = g.ntiles, p.tiles_per_seg = g.tiles_per_seg, p.nseg = g.nseg;
    p.PD = g.PD;
    p.nc = g.nc, p.nslots = g.nslots;
    const int64_t groups = (rows + kRbRows - 1) / kRbRows;
    if (groups * g.nseg * g.nc > INT32_MAX) return NBG_OK;
    auto launch = [&](auto kern) -> int {
        int r2 = allow_big_smem(kern, "nbg_group(rowbins2): cudaFuncSetAttribute");
        if (r2) return r2;
        kern<<<(unsigned)(groups * g.nseg * g.nc), g.NW * 32 + 32, g.smem, stream>>>(p);
        return check_launch("nbg_group(rowbins2)");
    };
    if (g.NW == 31)
        rc = g.runs ? launch(group_rowbins2_kernel<V, CLS, true, 31>) : launch(group_rowbins2_kernel<V, CLS, false, 31>);
    else
        rc = g.runs ? launch(group_rowbins2_kernel<V, CLS, true, 16>) : launch(group_rowbins2_kernel<V, CLS, false, 16>);
    if (rc) return rc;
    *handled = true;
    return NBG_OK;
}

}  // namespace nbg
