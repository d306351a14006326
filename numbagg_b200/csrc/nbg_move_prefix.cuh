// nbg_move_prefix.cuh -- moving windows of float32 data as differences of a streaming prefix.
//
// Replaces, for float32 inputs and windows > kDirectMax, the running-window form of
// move_rowtile_kernel (nbg_move.cu) on the loops of numbagg/moving.py:12-275.
//
// Why: on BASELINE config 4 (float32, window 1000) the running-window kernel forms every
// observation's widened contributions FOUR times (entering + leaving, in the re-sync pass and in
// the main pass): ~110 instructions per output, 8 float32->float64 conversions per element on the
// quarter-rate conversion pipe -- issue-bound at 0.18-0.32 of the HBM roofline (round-1 ncu).
// Here every observation is widened ONCE:
//   * a CTA streams a segment of one row tile by tile (2-stage TMA ring of raw input tiles);
//   * phase 1: each thread forms the contributions of its E consecutive observations
//     (value, products rounded to float32 first exactly as numba types them, valid flag), keeps
//     their running sums in registers, and one shuffle scan + one barrier turn them into the
//     segment-wide inclusive prefix P (double; exact for float32 data up to 2^29 observations of
//     one binade), which is written to a ring of the last window + TILE positions in shared memory;
//   * phase 2: output i = finalize(P[i] - P[i - window]) with P[i] still in registers and
//     P[i - window] one 16-byte shared-memory load per channel pair; results leave through a
//     double-buffered out tile and TMA bulk stores.
// Two CTA-wide barriers per tile; ~35 instructions per output for std (3 conversions).
// The prefix restarts with every segment (a few hundred thousand observations), which bounds its
// magnitude: the rounding error of a window sum is at most one ulp of the segment prefix, ~1e-11
// relative to the window sum at window = 1000 -- far inside the float32 tolerance (1e-5) and
// smaller than the drift of the reference's never re-synced running sums.  float64 inputs keep
// the running-window kernel: their tolerance (1e-12) is tighter than a prefix difference allows.
//
// Algorithmic traffic: one read per input element + one write per output element.
#pragma once

#include "nbg_common.cuh"

namespace nbg {

struct MovePfxParams {
    const void *a, *b;
    void *out;
    const void *a_halo, *b_halo;
    int64_t halo_len;
    int64_t rows, n;
    int window;
    int min_count;
    int segs_per_row;
    int tiles_per_seg;  // output tiles per segment
    int wup;            // window rounded up to a multiple of E (ring = wup + TILE positions)
};

// shared-memory carve-up (bytes), shared by host sizing and the kernel
template <int NIN, int NCH, int THREADS, int E>
struct PfxSmem {
    static constexpr int TILE = THREADS * E;
    static constexpr int NW = THREADS / 32;
    static constexpr int NPAIR = NCH / 2, NODD = NCH % 2;
    __host__ __device__ static size_t header() { return 64; }                                               // 2 mbarriers
    __host__ __device__ static size_t scratch() { return (size_t)(NW + 1) * (NCH + 1) * sizeof(double); }    // warp totals
    __host__ __device__ static size_t rcp(int w) { return ((size_t)(w + 3) * 8 + 15) & ~(size_t)15; }        // rcp[c + 1] = 1 / c
    __host__ __device__ static size_t in_stage() { return ((size_t)TILE * 4 + 32 + 15) & ~(size_t)15; }      // one raw input tile
    __host__ __device__ static size_t out_buf() { return ((size_t)TILE * 4 + 32 + 15) & ~(size_t)15; }
    __host__ __device__ static int ring_slots(int wup) { return wup + TILE + E; }                           // + mirrored first chunk
    __host__ __device__ static size_t ring(int wup) {
        return (size_t)ring_slots(wup) * (NPAIR * 16 + NODD * 8 + 4) + 64;
    }
    __host__ __device__ static size_t total(int w, int wup) {
        return header() + ((scratch() + 15) & ~(size_t)15) + rcp(w) + 2 * NIN * in_stage() + 2 * out_buf() + ring(wup);
    }
};

template <typename T, class Op, int THREADS, int E>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 || THREADS == 448) ? 2 : 1) move_prefix_kernel(MovePfxParams p) {
    static_assert(sizeof(T) == 4, "prefix windows are built for float32 inputs");
    constexpr int NIN = Op::NIN, NCH = Op::NCH, NCHP = NCH + 1;
    using SM = PfxSmem<NIN, NCH, THREADS, E>;
    constexpr int TILE = SM::TILE, NW = SM::NW, NPAIR = SM::NPAIR, NODD = SM::NODD;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t row = blockIdx.x / p.segs_per_row;
    const int seg = blockIdx.x % p.segs_per_row;
    const int w = p.window;
    const int H = (w + TILE - 1) / TILE;  // halo tiles: prefix only, no outputs
    const int64_t s0 = (int64_t)seg * p.tiles_per_seg * TILE;
    int64_t left = p.n - s0;
    const int out_tiles = (int)min((int64_t)p.tiles_per_seg, (left + TILE - 1) / TILE);
    const int ntl = H + out_tiles;
    const int64_t p_begin = s0 - (int64_t)H * TILE;
    const int R = p.wup + TILE;  // ring positions (multiple of E)

    // ---- carve shared memory
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);  // [2] input stages
    unsigned char *cur = smem_raw + SM::header();
    double *scratch = reinterpret_cast<double *>(cur);
    int *scratch_c = reinterpret_cast<int *>(scratch + NCH * (NW + 1));  // warp totals of the valid counts
    cur += (SM::scratch() + 15) & ~(size_t)15;
    double *rcp = reinterpret_cast<double *>(cur);  // rcp[c + 1] = 1 / c; rcp[0] = rcp[1] = 0 (count 0 / count - 1 == -1)
    cur += SM::rcp(w);
    unsigned char *in_base = cur;
    cur += 2 * NIN * SM::in_stage();
    unsigned char *out_base = cur;
    cur += 2 * SM::out_buf();
    const int slots = SM::ring_slots(p.wup);
    double2 *ringP = reinterpret_cast<double2 *>(cur);  // [NPAIR][slots]
    double *ringO = reinterpret_cast<double *>(cur + (size_t)NPAIR * slots * 16);  // [NODD][slots]
    int *ringC = reinterpret_cast<int *>(cur + (size_t)NPAIR * slots * 16 + (size_t)NODD * slots * 8);

    const T *row_a = reinterpret_cast<const T *>(p.a) + row * p.n;
    const T *row_b = NIN == 2 ? reinterpret_cast<const T *>(p.b) + row * p.n : nullptr;
    T *row_out = reinterpret_cast<T *>(p.out) + row * p.n;
    const T *halo_a = p.a_halo ? reinterpret_cast<const T *>(p.a_halo) + row * p.halo_len : nullptr;
    const T *halo_b = (NIN == 2 && p.b_halo) ? reinterpret_cast<const T *>(p.b_halo) + row * p.halo_len : nullptr;

    // TILE * 4 is a multiple of 16, so the 16-byte phase of a tile start is the same for every tile of
    // the segment: shared-memory placement and alignment tests are per-segment constants
    const uint32_t ph_a = span_phase(row_a, p_begin), ph_b = NIN == 2 ? span_phase(row_b, p_begin) : 0u;
    const uint32_t ph_o = span_phase(row_out, s0);
    const bool in_aligned = ph_a == 0 && ph_b == 0;  // whole interior tiles move with one bulk copy, no edges
    auto stage_ptr = [&](int st, int in) -> T * {
        return reinterpret_cast<T *>(in_base + (size_t)(st * NIN + in) * SM::in_stage() + (in ? ph_b : ph_a));
    };
    // stage `st` <- tile k: one bulk copy per input (thread 0) + thread-filled edges (everyone)
    auto issue_tile = [&](int k) {
        const int st = k & 1;
        const int64_t q0 = p_begin + (int64_t)k * TILE;
        T *sa = stage_ptr(st, 0);
        T *sb = NIN == 2 ? stage_ptr(st, 1) : nullptr;
        if (in_aligned && q0 >= 0 && q0 + TILE <= p.n) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar[st], (uint32_t)(NIN * TILE * 4));
                bulk_g2s(sa, row_a + q0, (uint32_t)(TILE * 4), &bar[st]);
                if (NIN == 2) bulk_g2s(sb, row_b + q0, (uint32_t)(TILE * 4), &bar[st]);
            }
            return;
        }
        const SpanPlan<T> pla = span_plan(row_a, q0, TILE, p.n);
        SpanPlan<T> plb = pla;
        if (NIN == 2) plb = span_plan(row_b, q0, TILE, p.n);
        if (tid == 0) {
            const uint32_t tx = pla.blk_bytes + (NIN == 2 ? plb.blk_bytes : 0u);
            mbar_arrive_expect_tx(&bar[st], tx);  // tx == 0: the arrival alone completes the phase
            if (pla.blk_bytes) bulk_g2s(sa + pla.blk_lo, row_a + q0 + pla.blk_lo, pla.blk_bytes, &bar[st]);
            if (NIN == 2 && plb.blk_bytes) bulk_g2s(sb + plb.blk_lo, row_b + q0 + plb.blk_lo, plb.blk_bytes, &bar[st]);
        }
        span_fill_edges<T, THREADS>(sa, row_a, q0, TILE, pla, quiet_nan<T>(), halo_a, p.halo_len);
        if (NIN == 2) span_fill_edges<T, THREADS>(sb, row_b, q0, TILE, plb, quiet_nan<T>(), halo_b, p.halo_len);
    };
    // outputs of tile k leave: one bulk store when the tile is whole and the phases agree
    auto store_tile = [&](int k) {
        const int64_t c0 = s0 + (int64_t)(k - H) * TILE;
        const int64_t rem = p.n - c0;
        const T *ob = reinterpret_cast<const T *>(out_base + (size_t)(k & 1) * SM::out_buf() + ph_o);
        if (ph_o == 0 && rem >= TILE) {
            if (tid == 0) {
                bulk_s2g(row_out + c0, ob, (uint32_t)(TILE * 4));
                bulk_commit();
            }
            return;
        }
        span_store<T, THREADS>(row_out + c0, ob, rem < TILE ? (int)rem : TILE);
        if (tid == 0 && (((uintptr_t)(row_out + c0)) & 15) != (smem_u32(ob) & 15)) bulk_commit();  // keep one group per tile
    };

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    for (int c = tid; c <= w + 1; c += THREADS) rcp[c] = c >= 2 ? 1.0 / (double)(c - 1) : 0.0;
    __syncthreads();
    issue_tile(0);
    if (ntl > 1) issue_tile(1);
    __syncthreads();  // edge fills of the first two tiles are visible

    double base[NCH];
#pragma unroll
    for (int q = 0; q < NCH; q++) base[q] = 0.0;
    int base_c = 0;
    int slot_w = tid * E;                              // ring slot of this thread's first position in tile 0
    int slot_r = ((tid * E - w) % R + R) % R;          // ... of (first position - window)
    const int mc = p.min_count;

    for (int k = 0; k < ntl; k++) {
        const int st = k & 1;
        mbar_wait(&bar[st], (uint32_t)((k >> 1) & 1));
        const T *sa = stage_ptr(st, 0);
        const T *sb = NIN == 2 ? stage_ptr(st, 1) : nullptr;

        // ---- phase 1: contributions once, running sums in registers
        double P[E][NCH];
        int C[E];
        {
            double acc[NCH];
#pragma unroll
            for (int q = 0; q < NCH; q++) acc[q] = 0.0;
            int cnt = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
                const T av = sa[tid * E + e];
                const T bv = NIN == 2 ? sb[tid * E + e] : av;
                const bool v = obs_valid<Op>(av, bv);
                double c[NCH];
                Op::contrib(v ? av : (T)0, v ? bv : (T)0, c);
#pragma unroll
                for (int q = 0; q < NCH; q++) {
                    acc[q] = dadd(acc[q], c[q]);
                    P[e][q] = acc[q];
                }
                cnt += v ? 1 : 0;
                C[e] = cnt;
            }
        }
        // inclusive scan of the thread totals across the warp
        double inc[NCH];
        int inc_c = C[E - 1];
#pragma unroll
        for (int q = 0; q < NCH; q++) inc[q] = P[E - 1][q];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                const double o = __shfl_up_sync(0xffffffffu, inc[q], d);
                if (lane >= d) inc[q] = dadd(inc[q], o);
            }
            const int oc = __shfl_up_sync(0xffffffffu, inc_c, d);
            if (lane >= d) inc_c += oc;
        }
        if (lane == 31) {
#pragma unroll
            for (int q = 0; q < NCH; q++) scratch[q * (NW + 1) + wid] = inc[q];
            scratch_c[wid] = inc_c;
        }
        // exclusive prefix of this lane inside the warp
        double off[NCH];
        int off_c;
#pragma unroll
        for (int q = 0; q < NCH; q++) {
            const double o = __shfl_up_sync(0xffffffffu, inc[q], 1);
            off[q] = lane ? o : 0.0;
        }
        {
            const int oc = __shfl_up_sync(0xffffffffu, inc_c, 1);
            off_c = lane ? oc : 0;
        }
        __syncthreads();  // B1: warp totals published; everyone is done with stage st and with tile k-1's phase 2

        // the stage just consumed takes tile k+2; tile k-1's outputs leave
        if (k + 2 < ntl) issue_tile(k + 2);
        if (k - 1 >= H) store_tile(k - 1);
        // prefix of the preceding warps + tile total: every warp scans the NW warp totals itself
        {
            double wt[NCH], winc[NCH];
#pragma unroll
            for (int q = 0; q < NCH; q++) winc[q] = wt[q] = lane < NW ? scratch[q * (NW + 1) + lane] : 0.0;
            const int wtc = lane < NW ? scratch_c[lane] : 0;
            int wincc = wtc;
#pragma unroll
            for (int d = 1; d < NW; d <<= 1) {
#pragma unroll
                for (int q = 0; q < NCH; q++) {
                    const double o = __shfl_up_sync(0xffffffffu, winc[q], d);
                    if (lane >= d) winc[q] = dadd(winc[q], o);
                }
                const int oc = __shfl_up_sync(0xffffffffu, wincc, d);
                if (lane >= d) wincc += oc;
            }
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                const double excl = __shfl_sync(0xffffffffu, winc[q] - wt[q], wid);  // exact: sums of float32 images
                const double tot = __shfl_sync(0xffffffffu, winc[q], NW - 1);
                off[q] = dadd(dadd(base[q], excl), off[q]);
                base[q] = dadd(base[q], tot);
            }
            off_c += base_c + __shfl_sync(0xffffffffu, wincc - wtc, wid);
            base_c += __shfl_sync(0xffffffffu, wincc, NW - 1);
        }
        // ---- absolute prefixes into the ring (and kept in registers for phase 2)
        {
            const bool mirror = slot_w == 0;  // the first chunk of the ring is mirrored behind its end
#pragma unroll
            for (int e = 0; e < E; e++) {
#pragma unroll
                for (int q = 0; q < NCH; q++) P[e][q] = dadd(P[e][q], off[q]);
                C[e] += off_c;
                const int sl = slot_w + e;
#pragma unroll
                for (int pr = 0; pr < NPAIR; pr++) ringP[pr * slots + sl] = make_double2(P[e][2 * pr], P[e][2 * pr + 1]);
                if (NODD) ringO[sl] = P[e][NCH - 1];
                ringC[sl] = C[e];
                if (mirror) {
#pragma unroll
                    for (int pr = 0; pr < NPAIR; pr++) ringP[pr * slots + R + e] = make_double2(P[e][2 * pr], P[e][2 * pr + 1]);
                    if (NODD) ringO[R + e] = P[e][NCH - 1];
                    ringC[R + e] = C[e];
                }
            }
        }
        if (tid == 0 && k >= H + 2) {
            // out buffer (k & 1) was handed to the TMA engine two tiles ago: its reads must be done
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncthreads();  // B2: ring complete

        // ---- phase 2: window sums as prefix differences, finalisation, out tile
        if (k >= H) {
            T *ob = reinterpret_cast<T *>(out_base + (size_t)(k & 1) * SM::out_buf() + ph_o);
            bool suspect = false;
            auto window_of = [&](int e, double *s) -> int {
                const int sl = slot_r + e;  // < R + E: the mirror makes the wrap contiguous
#pragma unroll
                for (int pr = 0; pr < NPAIR; pr++) {
                    const double2 m = ringP[pr * slots + sl];
                    s[2 * pr] = dsub(P[e][2 * pr], m.x);
                    s[2 * pr + 1] = dsub(P[e][2 * pr + 1], m.y);
                }
                if (NODD) s[NCH - 1] = dsub(P[e][NCH - 1], ringO[sl]);
                return C[e] - ringC[sl];
            };
#pragma unroll
            for (int e = 0; e < E; e++) {
                double s[NCH];
                const int count = window_of(e, s);
                const T r = Op::finalize_pfx(s, rcp[count + 1], rcp[count], suspect);
                ob[tid * E + e] = (count >= mc) ? r : quiet_nan<T>();
            }
            if (suspect) {
                // some output's float32 image left the normal range (zero / tiny / huge variance, the NaN
                // gate of a correlation): redo this thread's outputs with the exact finalisation
                for (int e = 0; e < E; e++) {
                    double s[NCH];
                    int count = 0;
#pragma unroll
                    for (int e2 = 0; e2 < E; e2++)
                        if (e2 == e) count = window_of(e2, s);
                    const T r = Op::finalize_fast(s, rcp[count + 1], rcp[count]);
                    ob[tid * E + e] = (count >= mc) ? r : quiet_nan<T>();
                }
            }
            fence_async_smem();
        }
        slot_w += TILE;
        if (slot_w >= R) slot_w -= R;
        slot_r += TILE;
        if (slot_r >= R) slot_r -= R;
    }
    __syncthreads();
    store_tile(ntl - 1);  // the last tile's outputs
    if (tid == 0) bulk_wait_read_all();
}

// Geometry per op: THREADS x E outputs per tile (E odd: conflict-free strided shared-memory access).
template <class Op, int G = 0>
struct PfxCfg {
    // the ring costs 12 / 20 / 28 / 44 bytes per position (mean / var / cov / corr): wider records get
    // smaller tiles so that window 1000 still fits one CTA per SM
    // (two CTAs of 256 threads per SM for mean / sum / var / std at window 1000: they hide each other's
    // two barriers per tile)
    static constexpr int THREADS = Op::NCH <= 2 ? 256 : 384;
    static constexpr int E = Op::NCH <= 2 ? 9 : (Op::NCH == 3 ? 7 : 5);
};
// G = 1 (experiment): 448 threads x 5 outputs, two CTAs = 28 warps per SM at window 1000 for the
// one- and two-channel ops (the 256 x 9 form is latency-bound at 16 warps: issue slots 48 % busy)
template <class Op>
struct PfxCfg<Op, 1> {
    static constexpr int THREADS = 448;
    static constexpr int E = 5;
};

template <typename T, class Op>
static bool prefix_fits(int64_t window) {
    using C = PfxCfg<Op>;
    using SM = PfxSmem<Op::NIN, Op::NCH, C::THREADS, C::E>;
    if (window > (1 << 20)) return false;
    const int wup = (int)((window + C::E - 1) / C::E * C::E);
    return SM::total((int)window, wup) <= kMaxSmemOptIn;
}

template <typename T, class Op, int G = 0>
static int launch_prefix(MovePfxParams p, int64_t outer, int64_t n, cudaStream_t stream) {
    using C = PfxCfg<Op, G>;
    if constexpr (G == 0 && Op::NCH <= 2) {
        if (const char *e = getenv("NBG_PFX_GEOM")) {
            using C1 = PfxCfg<Op, 1>;
            using SM1 = PfxSmem<Op::NIN, Op::NCH, C1::THREADS, C1::E>;
            const int wup1 = (p.window + C1::E - 1) / C1::E * C1::E;
            if (atoi(e) == 1 && SM1::total(p.window, wup1) <= kMaxSmemOptIn) return launch_prefix<T, Op, 1>(p, outer, n, stream);
        }
    }
    using SM = PfxSmem<Op::NIN, Op::NCH, C::THREADS, C::E>;
    p.wup = (p.window + C::E - 1) / C::E * C::E;
    const int64_t tpr = (n + SM::TILE - 1) / SM::TILE;
    // segments: enough CTAs for ~8 waves of the machine, but long enough to amortise the halo tiles
    const int H = (p.window + SM::TILE - 1) / SM::TILE;
    int64_t segs = (8 * (int64_t)kNumSMs + outer - 1) / outer;
    const int64_t max_segs = (tpr + 16 * H - 1) / (16 * H);  // >= 16 output tiles per halo tile
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    if (const char *e = getenv("NBG_PFX_SEGS")) segs = atoi(e) > 0 ? atoi(e) : segs;
    int64_t tps = (tpr + segs - 1) / segs;
    segs = (tpr + tps - 1) / tps;
    if (segs * outer > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_move: more than 2^31 segments");
    p.segs_per_row = (int)segs;
    p.tiles_per_seg = (int)tps;
    auto kern = move_prefix_kernel<T, Op, C::THREADS, C::E>;
    int rc = allow_big_smem(kern, "nbg_move(prefix): cudaFuncSetAttribute");
    if (rc) return rc;
    kern<<<(unsigned)(segs * outer), C::THREADS, SM::total(p.window, p.wup), stream>>>(p);
    return check_launch("nbg_move(prefix)");
}

}  // namespace nbg
