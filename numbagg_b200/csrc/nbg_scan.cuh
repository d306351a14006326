// nbg_scan.cuh -- single-pass chained scan along a row with decoupled look-back.
//
// Used by ffill/bfill (nbg_fill.cu) and the exponential moving functions (nbg_move_exp.cu):
// both are scans of an associative, NON-commutative aggregate along the core axis.
//
// Row tiles are taken in blockIdx order (see the kernel for the forward-progress argument).
// Per tile the CTA
//   1. stages the tile with one TMA bulk copy per input (nbg_common.cuh),
//   2. reduces each thread's E-element chunk from the identity state        (pass A),
//   3. scans the chunk aggregates across the CTA (warp shuffles + one smem round),
//   4. publishes the tile aggregate, looks back over predecessor tiles of the same row
//      (one warp, 32 descriptors per round) until it meets an inclusive prefix, publishes
//      its own inclusive prefix,
//   5. re-runs each chunk from its now-known incoming state and emits outputs (pass B),
//   6. drains the tile (outputs overwrite the staged input in place) with one bulk store.
// Each element is read from HBM once and written once: 2*itemsize algorithmic bytes.
//
// An `Agg` type provides:  static Agg identity();  static Agg combine(older, newer);
// static bool absorbing(a);  and is a POD of 8-byte words (shuffled / published word by word).
#pragma once

#include "nbg_common.cuh"

namespace nbg {

// ---- tile descriptors -------------------------------------------------------------------
// One descriptor per tile: the aggregate's 8-byte words, each stored as a 16-byte
// {word, tag} pair with a single 16-byte relaxed store and read back with a single 16-byte
// relaxed load (L1-bypassing), the same single-transaction idiom CUB's decoupled look-back
// uses for its {status, value} words.  Because every pair carries its own tag, a reader
// needs no fence: a descriptor is usable once all its pairs show the same non-zero tag
// (1 = tile aggregate, 2 = inclusive prefix; the inclusive prefix overwrites the aggregate
// in place, so a reader that catches a mix simply polls again).  The descriptor array is
// zeroed before every launch.
enum : unsigned long long { kTileEmpty = 0, kTileAggregate = 1, kTileInclusive = 2 };

__device__ __forceinline__ void st_pair(ulonglong2 *p, unsigned long long word, unsigned long long tag) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(word), "l"(tag) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_pair(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

template <class Agg>
struct TileDesc {
    static constexpr int W = sizeof(Agg) / 8;
    ulonglong2 pair[W];
};

template <class Agg>
__device__ __forceinline__ void desc_publish(TileDesc<Agg> *d, const Agg &a, unsigned long long tag) {
    constexpr int W = TileDesc<Agg>::W;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(&a);
#pragma unroll
    for (int i = 0; i < W; i++) st_pair(&d->pair[i], s[i], tag);
}
// Returns the descriptor's tag, or kTileEmpty when it is not (consistently) published yet.
template <class Agg>
__device__ __forceinline__ unsigned long long desc_try_read(const TileDesc<Agg> *d, Agg &out) {
    constexpr int W = TileDesc<Agg>::W;
    unsigned long long *o = reinterpret_cast<unsigned long long *>(&out);
    ulonglong2 v[W];
#pragma unroll
    for (int i = 0; i < W; i++) v[i] = ld_pair(&d->pair[i]);
    unsigned long long tag = v[0].y;
#pragma unroll
    for (int i = 0; i < W; i++) {
        o[i] = v[i].x;
        if (v[i].y != tag) tag = kTileEmpty;
    }
    return tag;
}

template <class Agg>
__device__ __forceinline__ Agg agg_shfl_up(const Agg &a, int delta) {
    constexpr int W = sizeof(Agg) / 8;
    Agg r;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(&a);
    unsigned long long *d = reinterpret_cast<unsigned long long *>(&r);
#pragma unroll
    for (int i = 0; i < W; i++) d[i] = __shfl_up_sync(0xffffffffu, s[i], delta);
    return r;
}
template <class Agg>
__device__ __forceinline__ Agg agg_shfl_down(const Agg &a, int delta) {
    constexpr int W = sizeof(Agg) / 8;
    Agg r;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(&a);
    unsigned long long *d = reinterpret_cast<unsigned long long *>(&r);
#pragma unroll
    for (int i = 0; i < W; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
    return r;
}
template <class Agg>
__device__ __forceinline__ Agg agg_shfl(const Agg &a, int src) {
    constexpr int W = sizeof(Agg) / 8;
    Agg r;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(&a);
    unsigned long long *d = reinterpret_cast<unsigned long long *>(&r);
#pragma unroll
    for (int i = 0; i < W; i++) d[i] = __shfl_sync(0xffffffffu, s[i], src);
    return r;
}

template <class Agg>
struct ScanWorkspace {
    TileDesc<Agg> *desc;
    __host__ __device__ static size_t bytes(int64_t ntiles) { return (size_t)ntiles * sizeof(TileDesc<Agg>) + 256; }
    __host__ __device__ static ScanWorkspace carve(void *base) {
        ScanWorkspace w;
        w.desc = reinterpret_cast<TileDesc<Agg> *>(base);
        return w;
    }
};

// Exclusive scan of one Agg per thread across the CTA (thread order = sequence order).
// Returns the exclusive prefix for this thread; *tile_total receives the CTA aggregate.
// `scratch` holds THREADS/32 Aggs in shared memory.
template <class Agg, int THREADS>
__device__ __forceinline__ Agg block_scan_agg(const Agg &mine, Agg *scratch, Agg *tile_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    Agg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Agg o = agg_shfl_up(inc, d);
        if (lane >= d) inc = Agg::combine(o, inc);
    }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    Agg warp_prefix = Agg::identity();
    Agg total = Agg::identity();
#pragma unroll
    for (int w = 0; w < NW; w++) {
        Agg t = scratch[w];
        if (w < wid) warp_prefix = Agg::combine(warp_prefix, t);
        total = Agg::combine(total, t);
    }
    *tile_total = total;
    Agg lane_excl = agg_shfl_up(inc, 1);
    if (lane == 0) lane_excl = Agg::identity();
    return Agg::combine(warp_prefix, lane_excl);
}

// Decoupled look-back, executed by warp 0 only (all 32 lanes).  tile_lin: this tile's
// linear id; row_first: linear id of the first tile of the same row (which always publishes
// an inclusive prefix without looking back).  Each round inspects 32 predecessors; the walk
// stops at the nearest predecessor that carries an inclusive prefix or whose aggregate is
// ABSORBING (Agg::absorbing(a): combine(x, a) == a for every x -- e.g. a fill tile that
// contains a valid value, or an exp tile whose decay product underflowed to exactly 0).
// Returns the exclusive prefix (state entering this tile), identical in every lane.
template <class Agg>
__device__ __forceinline__ Agg lookback_exclusive(const ScanWorkspace<Agg> &ws, int64_t tile_lin, int64_t row_first) {
    const int lane = threadIdx.x & 31;
    Agg excl = Agg::identity();
    int64_t newest = tile_lin - 1;
    while (true) {
        const int64_t mine = newest - lane;
        const bool in_range = mine >= row_first;
        unsigned long long tag = kTileEmpty;
        Agg p = Agg::identity();
        if (in_range) {
            do {
                tag = desc_try_read(ws.desc + mine, p);
            } while (tag == kTileEmpty);
        }
        const unsigned stop_mask =
            __ballot_sync(0xffffffffu, in_range && (tag == kTileInclusive || Agg::absorbing(p)));
        // lanes [0, last] take part: up to and including the nearest stopping tile
        const int last = stop_mask ? (__ffs(stop_mask) - 1) : 31;
        if (!in_range || lane > last) p = Agg::identity();
        // ordered reduction: afterwards lane l holds p[l+..] (older) combined before p[l]
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Agg o = agg_shfl_down(p, d);
            if (lane + d < 32) p = Agg::combine(o, p);
        }
        const Agg window = agg_shfl(p, 0);
        excl = Agg::combine(window, excl);
        if (stop_mask) break;
        // no single tile stopped the walk, but the COMPOSITION walked so far may already be
        // absorbing (e.g. the decay product underflowed to 0 across several exp tiles)
        if (Agg::absorbing(excl)) break;
        newest -= 32;
        if (newest < row_first) break;  // cannot happen: the row's first tile is inclusive
    }
    return excl;
}

// ======================================================================================
// Generic row-tile chained-scan kernel.  Policy P supplies:
//   using T;  using Agg;  static constexpr int NSTREAM;   (input streams staged per tile)
//   static constexpr bool REV;                            (scan from the row end: bfill)
//   static constexpr bool OVERLAP_INDEPENDENT;            (chunks with an absorbing in-tile prefix
//                                                          skip the wait for the tile carry)
//   __device__ static const T* stream_row(const ScanParams&, int s, int64_t row);
//   __device__ static Agg load_carry(const ScanParams&, int64_t row);
//   __device__ static void store_agg(const ScanParams&, int64_t row, const Agg&);
//   template<int E, class Get> __device__ static Agg reduce(const ScanParams&, Get get, int cnt);
//   template<int E, class Get, class Put> __device__ static void scan(const ScanParams&, Agg state,
//                                                                     Get get, Put put, int cnt);
//   where get(s, k) is stream s at chunk-local position k and put(k, v) stores output k;
//   cnt = number of in-range positions of this thread's chunk (0..E).
// ======================================================================================
struct ScanParams {
    const void *in[3];
    void *out;
    int64_t rows, n;
    int tiles_per_row;
    const void *carry_in;
    void *agg_out;
    void *ws_base;
    int64_t ntiles;
    double alpha_scalar;
    double min_weight;
    int64_t limit;
    int alpha_nd;  // alpha stream: 0 = 1-D shared by rows, 1 = per row
    int prefetch_dist;  // tiles ahead whose input spans this CTA prefetches into L2 (0: off)
};

template <class P, int THREADS, int E>
struct ScanSmem {
    using T = typename P::T;
    using Agg = typename P::Agg;
    static constexpr int TILE = THREADS * E;
    static constexpr size_t header = 64 + sizeof(Agg) * (THREADS / 32 + 2);
    __host__ __device__ static size_t header_bytes() { return (header + 15) & ~(size_t)15; }
    __host__ __device__ static size_t stream_bytes() { return ((size_t)TILE * sizeof(T) + 16 + 15) & ~(size_t)15; }
    // outputs are written in place over stream 0 (each thread only rewrites its own chunk)
    __host__ __device__ static size_t total() { return header_bytes() + P::NSTREAM * stream_bytes(); }
};

template <class P, int THREADS, int E>
__global__ void __launch_bounds__(THREADS, P::MIN_CTAS) scan_rowtile_kernel(ScanParams p) {
    using T = typename P::T;
    using Agg = typename P::Agg;
    using SM = ScanSmem<P, THREADS, E>;
    constexpr int TILE = THREADS * E;
    constexpr int NS = P::NSTREAM;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    volatile int *s_flag = reinterpret_cast<volatile int *>(smem_raw + 16);  // carry published
    Agg *scratch = reinterpret_cast<Agg *>(smem_raw + 64);  // [THREADS/32] + carry + spare
    Agg *s_carry = scratch + THREADS / 32;
    unsigned char *streams = smem_raw + SM::header_bytes();

    const int tid = threadIdx.x;
    const ScanWorkspace<Agg> ws = ScanWorkspace<Agg>::carve(p.ws_base);
    // Tiles are taken in blockIdx order: CTAs of a 1-D grid are dispatched in increasing
    // blockIdx, so every predecessor tile is owned by a CTA that is already resident (the
    // same forward-progress assumption CUB's single-pass scan makes).
    const int64_t tile_lin = blockIdx.x;
    if (tid == 0) {
        *s_flag = 0;
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t row = tile_lin / p.tiles_per_row;
    const int tile = (int)(tile_lin % p.tiles_per_row);
    const int64_t c0 = (int64_t)tile * TILE;          // logical (scan-order) start
    const int64_t p0 = P::REV ? (p.n - c0 - TILE) : c0;  // physical start of the span

    // ---- stage inputs
    const T *rows_[NS];
    T *s_[NS];
    SpanPlan<T> pl[NS];
    uint32_t tx = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        rows_[s] = P::stream_row(p, s, row);
        s_[s] = reinterpret_cast<T *>(streams + s * SM::stream_bytes() + span_phase(rows_[s], p0));
        pl[s] = span_plan(rows_[s], p0, TILE, p.n);
        tx += pl[s].blk_bytes;
    }
    if (tid == 0 && tx > 0) {
        mbar_arrive_expect_tx(bar, tx);
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (pl[s].blk_bytes) bulk_g2s(s_[s] + pl[s].blk_lo, rows_[s] + p0 + pl[s].blk_lo, pl[s].blk_bytes, bar);
    }
    if (tid == 0 && p.prefetch_dist > 0) {
        // keep HBM busy for the wave of CTAs that replaces this one: its TMA loads then hit L2
        const int64_t tl = tile_lin + p.prefetch_dist;
        if (tl < p.ntiles) {
            const int64_t row2 = tl / p.tiles_per_row;
            const int64_t c2 = (tl % p.tiles_per_row) * (int64_t)TILE;
            const int64_t q0 = P::REV ? (p.n - c2 - TILE) : c2;
#pragma unroll
            for (int s = 0; s < NS; s++) span_prefetch_l2(P::stream_row(p, s, row2), q0, TILE, p.n);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++)
        span_fill_edges<T, THREADS>(s_[s], rows_[s], p0, TILE, pl[s], quiet_nan<T>(), (const T *)nullptr, 0);
    if (tx > 0) mbar_wait(bar, 0);
    __syncthreads();

    // ---- pass A: chunk aggregates from the identity state
    const int k0 = tid * E;
    const int64_t remaining = p.n - c0 - k0;
    const int cnt = remaining >= E ? E : (remaining > 0 ? (int)remaining : 0);
    auto get = [&](int s, int k) -> T {
        const int j = P::REV ? (TILE - 1 - (k0 + k)) : (k0 + k);
        return s_[s][j];
    };
    const Agg mine = P::template reduce<E>(p, get, cnt);
    Agg tile_total;
    const Agg excl = block_scan_agg<Agg, THREADS>(mine, scratch, &tile_total);

    // ---- tile carry: first tile of a row takes carry_in, the others look back.  Only
    // threads whose chunk state DEPENDS on the carry wait for it: a chunk whose in-tile
    // prefix is absorbing (a valid value earlier in the tile for fills, an underflowed decay
    // product for exp) starts pass B immediately, overlapping the look-back latency.
    const bool independent = P::OVERLAP_INDEPENDENT && Agg::absorbing(excl);
    if (tid < 32) {
        const int64_t row_first = row * p.tiles_per_row;
        Agg carry;
        if (tile == 0) {
            carry = p.carry_in ? P::load_carry(p, row) : Agg::identity();
        } else {
            if (tid == 0) desc_publish(ws.desc + tile_lin, tile_total, kTileAggregate);
            __syncwarp();
            carry = lookback_exclusive(ws, tile_lin, row_first);
        }
        if (tid == 0) {
            const Agg incl = Agg::combine(carry, tile_total);
            desc_publish(ws.desc + tile_lin, incl, kTileInclusive);
            *s_carry = carry;
            __threadfence_block();
            *s_flag = 1;
            if (p.agg_out && tile == p.tiles_per_row - 1) P::store_agg(p, row, incl);
        }
    }
    if (p.out == nullptr) return;

    // ---- pass B: re-run every chunk from its incoming state, emit outputs
    Agg state = excl;
    if (P::OVERLAP_INDEPENDENT) {
        if (!independent) {
            while (*s_flag == 0) {
            }
            __threadfence_block();
            state = Agg::combine(*s_carry, excl);
        }
    } else {
        __syncthreads();  // hardware barrier: cheaper than spinning when every chunk needs the carry
        state = Agg::combine(*s_carry, excl);
    }
    T *row_out = reinterpret_cast<T *>(p.out) + row * p.n;
    T *sout = s_[0];  // in place: put(k) overwrites the element get(0, k) already consumed
    auto put = [&](int k, T v) {
        const int j = P::REV ? (TILE - 1 - (k0 + k)) : (k0 + k);
        sout[j] = v;
    };
    P::template scan<E>(p, state, get, put, cnt);
    fence_async_smem();
    __syncthreads();
    // physical range of this tile clipped to the row
    const int64_t lo = p0 < 0 ? 0 : p0;
    const int64_t hi = (p0 + TILE > p.n) ? p.n : (p0 + TILE);
    if (hi > lo) span_store<T, THREADS>(row_out + lo, sout + (lo - p0), (int)(hi - lo));
    if (tid == 0) bulk_wait_read_all();
}

template <class P, int THREADS, int E>
static int launch_scan_rowtile(ScanParams p, int64_t rows, int64_t n, void *workspace, size_t workspace_bytes,
                               cudaStream_t stream, const char *what) {
    using SM = ScanSmem<P, THREADS, E>;
    using Agg = typename P::Agg;
    const int64_t tpr = (n + SM::TILE - 1) / SM::TILE;
    const int64_t ntiles = tpr * rows;
    if (ntiles > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "scan: more than 2^31 tiles");
    const size_t need = ScanWorkspace<Agg>::bytes(ntiles);
    if (workspace == nullptr || workspace_bytes < need) return fail(NBG_ERR_WORKSPACE, "scan: workspace too small");
    // align the workspace base to 128 bytes (the caller's allocation may be byte-aligned)
    uintptr_t base = ((uintptr_t)workspace + 127) & ~(uintptr_t)127;
    p.ws_base = reinterpret_cast<void *>(base);
    p.ntiles = ntiles;
    p.tiles_per_row = (int)tpr;
    p.rows = rows;
    p.n = n;
    p.prefetch_dist = prefetch_distance(P::MIN_CTAS, (size_t)P::NSTREAM * THREADS * E * sizeof(typename P::T));
    int rc = check_cuda(cudaMemsetAsync(p.ws_base, 0, (size_t)ntiles * sizeof(TileDesc<Agg>), stream), what);
    if (rc) return rc;
    auto kern = scan_rowtile_kernel<P, THREADS, E>;
    rc = allow_big_smem(kern, what);
    if (rc) return rc;
    kern<<<(unsigned)ntiles, THREADS, SM::total(), stream>>>(p);
    return check_launch(what);
}

template <class P, int THREADS, int E>
static size_t scan_rowtile_workspace_bytes(int64_t rows, int64_t n) {
    using SM = ScanSmem<P, THREADS, E>;
    const int64_t tpr = (n + SM::TILE - 1) / SM::TILE;
    return ScanWorkspace<typename P::Agg>::bytes(tpr * rows) + 128;
}

}  // namespace nbg
