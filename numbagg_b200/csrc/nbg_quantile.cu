// nbg_quantile.cu -- nanquantile / nanmedian over the rows of a (rows, n) float64 matrix
// (numbagg/funcs.py:245-291, 332-335 behind ndquantile, numbagg/decorators.py:821-901).
//
// The reference partitions a NaN-filled copy of every slice (np.partition) and interpolates
// linearly between the two order statistics around rank (valid - 1) * q.  Selection is exact,
// so only the final interpolation does arithmetic -- done here with the reference's own three
// roundings -- and results are bit-identical.  Values travel as order-preserving 64-bit keys
// (order_key; NaN = the largest key, so missing values sort last like the reference's fill
// with the maximum).
//   rows of <= 512 elements: one CTA loads the row into shared memory, sorts the keys with a
//     bitonic network and reads the order statistics off (quant_sort_kernel);
//   rows of <= 4096 elements: the radix select below, run by one CTA on the row's keys in
//     shared memory (quant_smem_select_kernel);
//   longer rows: radix select, 11 bits per pass from the top (shared-memory histograms per distinct
//     target prefix, flushed to global; a warp per target picks the bucket holding its rank);
//     after two passes the elements still matching a 22-bit prefix are compacted into a small
//     per-row candidate list, which one CTA sorts: 3 reads of the data (rows with more ties than
//     the list holds go on with histogram passes), any number of rows per launch.
#include "nbg_common.cuh"

namespace nbg {
namespace {

typedef unsigned long long u64;
typedef long long i64;

constexpr int kQThreads = 256;
constexpr int kQSortMax = 4096;  // longest row sorted in shared memory
constexpr int kQMaxQ = 16;       // quantiles per call (two targets each: floor and ceil rank)

__device__ __forceinline__ u64 quant_key(double x) { return x != x ? ~0ull : order_key(x); }

// funcs.py:268-288: rank = (valid - 1) * q; floor/ceil indexes; proportion = rank - floor;
// floor_val + proportion * (ceil_val - floor_val)
__device__ __forceinline__ void quant_rank(i64 valid, double q, double &rank, i64 &lo, i64 &hi) {
    rank = __dmul_rn((double)(valid - 1), q);
    lo = __double2ll_rd(rank);
    hi = __double2ll_ru(rank);
}
__device__ __forceinline__ double quant_interpolate(double fv, double cv, double rank, i64 lo) {
    const double proportion = __dsub_rn(rank, (double)lo);
    return __dadd_rn(fv, __dmul_rn(proportion, __dsub_rn(cv, fv)));
}

// --------------------------------------------------------------------- short rows: sort
__global__ void __launch_bounds__(kQThreads) quant_sort_kernel(const double *__restrict__ a,
                                                               const double *__restrict__ q,
                                                               double *__restrict__ out, i64 n, int npad, int m) {
    extern __shared__ __align__(16) unsigned char quant_smem[];
    u64 *keys = reinterpret_cast<u64 *>(quant_smem);
    __shared__ int s_valid;
    const int tid = threadIdx.x, T = blockDim.x;
    const i64 row = blockIdx.x;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < npad; i += T) {
        u64 k = ~0ull;
        if (i < n) {
            const double x = a[row * n + i];
            k = quant_key(x);
            local += x == x ? 1 : 0;
        }
        keys[i] = k;
    }
    for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((tid & 31) == 0 && local) atomicAdd(&s_valid, local);
    __syncthreads();
    // bitonic sort, ascending
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // bit log2(j) clear
                const int p = i | j;
                const bool up = (i & k) == 0;
                const u64 x = keys[i], y = keys[p];
                if ((x > y) == up) {
                    keys[i] = y;
                    keys[p] = x;
                }
            }
            __syncthreads();
        }
    }
    const i64 valid = s_valid;
    for (int t = tid; t < m; t += T) {
        const double qq = q[t];
        double r = quiet_nan<double>();
        if (valid > 0 && qq == qq) {
            double rank;
            i64 lo, hi;
            quant_rank(valid, qq, rank, lo, hi);
            r = quant_interpolate(key_to_double(keys[lo]), key_to_double(keys[hi]), rank, lo);
        }
        out[row * m + t] = r;
    }
}

// ------------------------------------------------- rows of <= 1024 elements: one WARP per row
// The row's keys live in registers (KPL per lane, blocked: element index = lane * KPL + r) and are
// sorted by a bitonic network: exchanges at distance < KPL stay inside a lane (two compares + four
// selects per pair), larger distances are one shuffle pair per key.  No shared memory, no barriers:
// at n = 1000 the CTA-per-row kernels below spend their time in ~130 CTA-wide barriers per row (31 ms
// on 10^6 x 1000); this one issues ~10 000 instructions per lane and row whatever the number of
// quantiles and runs at the limit of the integer pipe (17.6 ms; ncu `profiles/r02_ncu_quant_quartiles_short.txt`).
// Only the exchanges inside a lane need compile-time register indices; the block loop (kk) and the
// cross-lane distances (lj) are RUNTIME loops -- fully unrolled the network is 10 000 instructions
// (160 KB of code) and the four warps of a scheduler, each somewhere else in it, stalled on
// instruction fetch 70 % of the time (`no_instruction` 7.3 per issue, 31 ms).  Tried: sorting the
// values as doubles with fmin / fmax (they are not single instructions here: 21 000 per row, 27.6 ms).
// Loads are coalesced (the initial order of the keys is irrelevant to a sort).
template <int KPL>
__device__ __forceinline__ u64 key_at(const u64 (&k)[KPL], int r) {
    u64 v = k[0];
#pragma unroll
    for (int i = 1; i < KPL; i++) v = r == i ? k[i] : v;
    return v;
}

template <int KPL>
__global__ void __launch_bounds__(128) quant_warp_sort_kernel(const double *__restrict__ a, const double *__restrict__ q,
                                                              double *__restrict__ out, i64 rows, i64 n, int m) {
    constexpr int N = 32 * KPL;
    const int lane = threadIdx.x & 31;
    const i64 row = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;  // whole warps leave together
    const double *p = a + row * n;
    u64 k[KPL];
    int valid = 0;
#pragma unroll
    for (int r = 0; r < KPL; r++) {
        const int i = r * 32 + lane;
        u64 key = ~0ull;
        if (i < n) {
            const double x = __ldcs(p + i);
            key = quant_key(x);
            valid += x == x ? 1 : 0;
        }
        k[r] = key;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, d);
    // blocks shorter than a lane's share: direction and partners are register bits
#pragma unroll
    for (int kk = 2; kk < KPL; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < KPL; r++) {
                if ((r & j) == 0) {
                    const int r2 = r | j;
                    const bool up = (r & kk) == 0;
                    const u64 x = k[r], y = k[r2];
                    const bool sw = (x > y) == up;
                    k[r] = sw ? y : x;
                    k[r2] = sw ? x : y;
                }
            }
        }
    }
#pragma unroll 1
    for (int kk = KPL; kk <= N; kk <<= 1) {
        const bool up = ((lane * KPL) & kk) == 0;  // the block's direction is a lane bit from here on
        // partner in lane ^ lj, same register; this lane keeps the smaller key iff it is the lower index of
        // the pair in an ascending block (or the upper one in a descending block)
#pragma unroll 1
        for (int lj = kk / (2 * KPL); lj > 0; lj >>= 1) {
            const bool keep_min = ((lane & lj) == 0) == up;
#pragma unroll
            for (int r = 0; r < KPL; r++) {
                const u64 mine = k[r];
                const u64 other = __shfl_xor_sync(0xffffffffu, mine, lj);
                const bool take_other = keep_min ? (other < mine) : (other > mine);
                k[r] = take_other ? other : mine;
            }
        }
#pragma unroll
        for (int j = KPL >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < KPL; r++) {
                if ((r & j) == 0) {
                    const int r2 = r | j;
                    const u64 x = k[r], y = k[r2];
                    const bool sw = (x > y) == up;
                    k[r] = sw ? y : x;
                    k[r2] = sw ? x : y;
                }
            }
        }
    }
    for (int t = 0; t < m; t++) {
        const double qq = q[t];
        double res = quiet_nan<double>();
        if (valid > 0 && qq == qq) {  // uniform across the warp
            double rank;
            i64 lo, hi;
            quant_rank((i64)valid, qq, rank, lo, hi);
            const u64 mine_lo = key_at<KPL>(k, (int)(lo % KPL)), mine_hi = key_at<KPL>(k, (int)(hi % KPL));
            const u64 klo = __shfl_sync(0xffffffffu, mine_lo, (int)(lo / KPL));
            const u64 khi = __shfl_sync(0xffffffffu, mine_hi, (int)(hi / KPL));
            res = quant_interpolate(key_to_double(klo), key_to_double(khi), rank, lo);
        }
        if (lane == 0) out[row * m + t] = res;
    }
}

// ------------------------------------------------- medium rows: radix select in shared memory
// Same selection as the long-row path below, but the row's keys sit in shared memory, so the
// eight passes cost no DRAM traffic: ~10 instructions per element and pass, against ~900 per
// element for the bitonic network at n = 1000.  One CTA per row.
__global__ void __launch_bounds__(kQThreads) quant_smem_select_kernel(const double *__restrict__ a,
                                                                      const double *__restrict__ q,
                                                                      double *__restrict__ out, i64 n, int m) {
    extern __shared__ __align__(16) unsigned char quant_smem[];
    u64 *keys = reinterpret_cast<u64 *>(quant_smem);                               // [n]
    unsigned *h = reinterpret_cast<unsigned *>(quant_smem + (((size_t)n * 8 + 15) & ~(size_t)15));  // [T2][256]
    __shared__ u64 s_prefix[2 * kQMaxQ];
    __shared__ i64 s_rem[2 * kQMaxQ];
    __shared__ int s_alias[2 * kQMaxQ];
    __shared__ int s_counted[2 * kQMaxQ];
    __shared__ int s_tab[2 * kQMaxQ];
    __shared__ u64 s_pre[2 * kQMaxQ];
    __shared__ int s_ncounted;
    __shared__ int s_valid;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = blockDim.x;  // 128 for rows up to 2048 elements (more resident CTAs, cheaper barriers), else 256
    const int T2 = 2 * m;
    const i64 row = blockIdx.x;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < n; i += T) {
        const double x = a[row * n + i];
        keys[i] = quant_key(x);
        local += x == x ? 1 : 0;
    }
    for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if (lane == 0 && local) atomicAdd(&s_valid, local);
    __syncthreads();
    const i64 valid = s_valid;
    if (tid < T2) {
        const double qq = q[tid >> 1];
        i64 rem = -1;
        if (valid > 0 && qq == qq) {
            double rank;
            i64 lo, hi;
            quant_rank(valid, qq, rank, lo, hi);
            rem = (tid & 1) ? hi : lo;
        }
        s_rem[tid] = rem;
        s_prefix[tid] = 0;
    }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        // one histogram per distinct prefix among the active targets
        if (tid < T2) {
            int first = tid;
            for (int t = tid - 1; t >= 0; t--)
                if (s_rem[t] >= 0 && s_prefix[t] == s_prefix[tid]) first = t;
            s_alias[tid] = first;
            s_counted[tid] = (s_rem[tid] >= 0 && first == tid) ? 1 : 0;
        }
        __syncthreads();
        if (tid == 0) {  // compact (prefix, table) list of the histograms to fill
            int c = 0;
            for (int t = 0; t < T2; t++)
                if (s_counted[t]) {
                    s_tab[c] = t;
                    s_pre[c] = s_prefix[t];
                    c++;
                }
            s_ncounted = c;
        }
        for (int t = 0; t < T2; t++)  // only the tables that will be counted into
            if (s_counted[t])
                for (int i = tid; i < 256; i += T) h[t * 256 + i] = 0;
        __syncthreads();
        const int ncounted = s_ncounted;
        for (int i = tid; i < n; i += T) {
            const u64 key = keys[i];
            const unsigned digit = (unsigned)(key >> shift) & 255u;
            const u64 head = pass == 0 ? 0 : key >> (shift + 8);
            for (int c = 0; c < ncounted; c++)
                if (head == s_pre[c]) atomicAdd(&h[s_tab[c] * 256 + digit], 1u);
        }
        __syncthreads();
        // a warp per target: lane l owns bins [8l, 8l + 8); find the bin holding the rank
        for (int t = wid; t < T2; t += T / 32) {
            const i64 rem = s_rem[t];
            if (rem < 0) continue;  // uniform across the warp
            const unsigned *ht = h + s_alias[t] * 256 + lane * 8;
            unsigned c[8];
            unsigned mine = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                c[b] = ht[b];
                mine += c[b];
            }
            unsigned inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            const unsigned before = inc - mine;
            const bool here = rem >= (i64)before && rem < (i64)inc;
            const unsigned who = __ballot_sync(0xffffffffu, here);
            // `who` has exactly one bit (rem < total count of the bucket); the last lane
            // takes over if the counts were ever inconsistent
            const int owner = who ? __ffs(who) - 1 : 31;
            if (lane == owner) {
                i64 r2 = rem - (i64)before;
                int b = 0;
                for (; b < 7; b++) {
                    if (r2 < (i64)c[b]) break;
                    r2 -= (i64)c[b];
                }
                s_rem[t] = r2;
                s_prefix[t] = (s_prefix[t] << 8) | (u64)(lane * 8 + b);
            }
        }
        __syncthreads();
    }
    for (int t = tid; t < m; t += T) {
        const double qq = q[t];
        double r = quiet_nan<double>();
        if (valid > 0 && qq == qq) {
            double rank;
            i64 lo, hi;
            quant_rank(valid, qq, rank, lo, hi);
            r = quant_interpolate(key_to_double(s_prefix[2 * t]), key_to_double(s_prefix[2 * t + 1]), rank, lo);
        }
        out[row * m + t] = r;
    }
}

// ------------------------------------------------------------------ long rows: radix select
// 11-bit digits from the top (6 passes would finish the key: 5 x 11 + 9 bits), but after TWO
// histogram passes a target's bucket is 2^-22 of the key space -- a few hundred elements of a
// million -- so the third read of the data only COMPACTS the elements that still match some
// target's 22-bit prefix into a per-row candidate list (<= kQCap keys), and one CTA per row sorts
// that list in shared memory and reads every order statistic off it.  3 reads of the data instead
// of 8.  A row whose buckets hold more than kQCap candidates (heavy ties: every copy of a repeated
// value shares all 64 bits) keeps going with histogram passes 2..5; rows are independent, so the
// kernel sequence is fixed and finished rows simply skip the tail.
// Workspace per row: valid count; per target (2 per quantile) the key prefix found so far, the
// rank that remains inside that prefix's bucket (-1: inactive target) and the size of that bucket.
constexpr int kQBins = 2048;
constexpr int kQPasses = 6;
constexpr int kQCap = 2048;    // candidate keys per row (16 KB of shared memory in the finishing CTA)
constexpr int kQChunk = 12;    // quantiles per internal round: 24 histograms x 8 KB of shared memory
__host__ __device__ inline int q_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
__host__ __device__ inline int q_bits(int pass) { return pass < 5 ? 11 : 9; }

struct QuantWs {
    i64 *valid;         // [rows]
    u64 *prefix;        // [rows][T2]
    i64 *remaining;     // [rows][T2]
    unsigned *bucket;   // [rows][T2]: elements in the bucket chosen by the last select
    int *alias;         // [rows][T2]: first target with the same prefix (its histogram serves both)
    unsigned *hist;     // [rows][T2][kQBins]
    int *done;          // [rows]: 1 = finished from the candidate list
    unsigned *ncand;    // [rows]
    u64 *cand;          // [rows][kQCap]
};

// bin += 1 for every lane with `pred`, one shared-memory atomic per distinct digit in the warp
__device__ __forceinline__ void warp_hist_add(unsigned *bins, unsigned digit, bool pred) {
    const unsigned active = __ballot_sync(0xffffffffu, pred);
    if (!pred) return;
    const unsigned peers = __match_any_sync(active, digit);
    if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&bins[digit], (unsigned)__popc(peers));
}

// The distinct prefixes among a row's active targets, as a compact (prefix, owner target) list in
// shared memory; returns their number.  alias_out (may be null): first target with the same prefix.
__device__ __forceinline__ int quant_distinct(const QuantWs &ws, i64 row, int T2, u64 *s_pre, int *s_tab, unsigned *s_cnt,
                                              int *alias_out) {
    __shared__ u64 s_prefix[2 * kQMaxQ];
    __shared__ int s_active[2 * kQMaxQ];
    __shared__ int s_first[2 * kQMaxQ];
    __shared__ int s_n;
    const int tid = threadIdx.x;
    if (tid < T2) {
        s_prefix[tid] = ws.prefix[row * T2 + tid];
        s_active[tid] = ws.remaining[row * T2 + tid] >= 0 ? 1 : 0;
    }
    __syncthreads();
    if (tid < T2) {
        int first = tid;
        for (int t = tid - 1; t >= 0; t--)
            if (s_active[t] && s_prefix[t] == s_prefix[tid]) first = t;
        s_first[tid] = first;
        if (alias_out) alias_out[row * T2 + tid] = first;
    }
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        for (int t = 0; t < T2; t++)
            if (s_active[t] && s_first[t] == t) {
                s_tab[c] = t;
                s_pre[c] = s_prefix[t];
                if (s_cnt) s_cnt[c] = ws.bucket[row * T2 + t];
                c++;
            }
        s_n = c;
    }
    __syncthreads();
    return s_n;
}

// pass = 0: one histogram of the top digit per row (+ the count of non-NaN elements);
// pass >= 1: per distinct prefix, the next digit of the elements that match it.
// grid = (segments, rows)
__global__ void __launch_bounds__(kQThreads) quant_hist_kernel(const double *__restrict__ a, QuantWs ws, i64 n, int T2,
                                                               int pass) {
    extern __shared__ __align__(16) unsigned char quant_smem[];
    unsigned *h = reinterpret_cast<unsigned *>(quant_smem);  // [ntab][kQBins]
    __shared__ int s_tab[2 * kQMaxQ];
    __shared__ u64 s_pre[2 * kQMaxQ];
    __shared__ int s_valid;
    const int tid = threadIdx.x;
    const i64 row = blockIdx.y;
    if (pass >= 2 && ws.done[row]) return;
    int ncounted = 0;
    if (pass > 0) ncounted = quant_distinct(ws, row, T2, s_pre, s_tab, nullptr, blockIdx.x == 0 ? ws.alias : nullptr);
    const int nbins = 1 << q_bits(pass);
    if (pass == 0) {
        for (int i = tid; i < nbins; i += kQThreads) h[i] = 0;
    } else {
        for (int c = 0; c < ncounted; c++)
            for (int i = tid; i < nbins; i += kQThreads) h[s_tab[c] * kQBins + i] = 0;
    }
    if (tid == 0) s_valid = 0;
    __syncthreads();
    const u64 pre0 = s_pre[0];
    unsigned *h0 = h + (ncounted > 0 ? s_tab[0] : 0) * kQBins;
    const double *p = a + row * n;
    const int shift = q_shift(pass), hshift = shift + q_bits(pass);
    const unsigned mask = (unsigned)nbins - 1u;
    int local_valid = 0;
    // whole warps iterate together (the histogram update uses warp votes)
    const i64 per_cta = ((n + gridDim.x - 1) / gridDim.x + 31) & ~(i64)31;
    const i64 lo = (i64)blockIdx.x * per_cta;
    const i64 hi = lo + per_cta < n ? lo + per_cta : n;
    constexpr int U = 8;  // independent loads in flight per thread (64 bytes)
    for (i64 base = lo; base < hi; base += (i64)U * kQThreads) {
        double xs[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 i = base + (i64)u * kQThreads + tid;
            xs[u] = i < hi ? __ldcs(p + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool in = base + (i64)u * kQThreads + tid < hi;
            const double x = xs[u];
            const u64 key = quant_key(x);
            const unsigned digit = (unsigned)(key >> shift) & mask;
            if (pass == 0) {
                local_valid += (in && x == x) ? 1 : 0;
                // sign and exponent: a handful of distinct digits per warp -> warp-aggregated update
                // (tried: per-thread register counts for the two most recent digits -- divergent, 4.7 vs 4.4 ms)
                warp_hist_add(h, digit, in);
            } else {
                const u64 head = key >> hshift;
                if (ncounted == 1) {
                    if (in && head == pre0) atomicAdd(&h0[digit], 1u);
                } else {
                    for (int c = 0; c < ncounted; c++)
                        if (in && head == s_pre[c]) atomicAdd(&h[s_tab[c] * kQBins + digit], 1u);
                }
            }
        }
    }
    if (pass == 0) {
        for (int d = 16; d >= 1; d >>= 1) local_valid += __shfl_xor_sync(0xffffffffu, local_valid, d);
        if ((tid & 31) == 0 && local_valid) atomicAdd(&s_valid, local_valid);
    }
    __syncthreads();
    unsigned *g = ws.hist + (size_t)row * T2 * kQBins;
    if (pass == 0) {
        for (int i = tid; i < nbins; i += kQThreads)
            if (h[i]) atomicAdd(&g[i], h[i]);
    } else {
        for (int c = 0; c < ncounted; c++)
            for (int i = tid; i < nbins; i += kQThreads) {
                const int w = s_tab[c] * kQBins + i;
                if (h[w]) atomicAdd(&g[w], h[w]);
            }
    }
    if (pass == 0 && tid == 0 && s_valid) atomicAdd(reinterpret_cast<u64 *>(&ws.valid[row]), (u64)s_valid);
}

// one WARP per (row, target): pick the bucket that holds the remaining rank and extend the prefix.
// After pass 0 the targets are created first.
__global__ void __launch_bounds__(128) quant_select_kernel(QuantWs ws, const double *__restrict__ q, i64 rows, int T2, int pass) {
    const i64 gid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gid >= rows * T2) return;
    const i64 row = gid / T2;
    const int t = (int)(gid % T2);
    if (pass >= 2 && ws.done[row]) return;
    const unsigned *h = ws.hist + ((size_t)row * T2 + (pass == 0 ? 0 : ws.alias[gid])) * kQBins;
    i64 rem = ws.remaining[gid];
    u64 prefix = ws.prefix[gid];
    if (pass == 0) {
        const i64 valid = ws.valid[row];
        const double qq = q[t >> 1];
        rem = -1;
        prefix = 0;
        if (valid > 0 && qq == qq) {
            double rank;
            i64 lo, hi;
            quant_rank(valid, qq, rank, lo, hi);
            rem = (t & 1) ? hi : lo;
        }
    }
    unsigned bucket = 0;
    if (rem >= 0) {  // uniform across the warp
        const int nbins = 1 << q_bits(pass), per = nbins / 32;
        const unsigned *hl = h + lane * per;
        unsigned mine = 0;
        for (int b = 0; b < per; b++) mine += hl[b];
        unsigned inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const unsigned before = inc - mine;
        const bool here = rem >= (i64)before && rem < (i64)inc;
        const unsigned who = __ballot_sync(0xffffffffu, here);
        const int owner = who ? __ffs(who) - 1 : 31;  // (the last lane takes over if the counts were inconsistent)
        int d = 0;
        i64 r2 = 0;
        if (lane == owner) {
            r2 = rem - (i64)before;
            for (; d < per - 1; d++) {
                const i64 c = (i64)hl[d];
                if (r2 < c) break;
                r2 -= c;
            }
            bucket = hl[d];
            d += lane * per;
        }
        d = __shfl_sync(0xffffffffu, d, owner);
        r2 = __shfl_sync(0xffffffffu, r2, owner);
        bucket = __shfl_sync(0xffffffffu, bucket, owner);
        rem = r2;
        prefix = (prefix << q_bits(pass)) | (u64)d;
    }
    if (lane == 0) {
        ws.remaining[gid] = rem;
        ws.prefix[gid] = prefix;
        ws.bucket[gid] = bucket;
    }
}

// Third read of the data (after passes 0 and 1): elements that match a target's 22-bit prefix go to
// the row's candidate list -- if the buckets of the row's distinct prefixes fit in it.
__global__ void __launch_bounds__(kQThreads) quant_compact_kernel(const double *__restrict__ a, QuantWs ws, i64 n, int T2) {
    __shared__ int s_tab[2 * kQMaxQ];
    __shared__ u64 s_pre[2 * kQMaxQ];
    __shared__ unsigned s_cnt[2 * kQMaxQ];
    const int tid = threadIdx.x;
    const i64 row = blockIdx.y;
    const int nd = quant_distinct(ws, row, T2, s_pre, s_tab, s_cnt, nullptr);
    unsigned total = 0;
    for (int c = 0; c < nd; c++) total += s_cnt[c];
    if (nd == 0 || total > (unsigned)kQCap) return;  // no active target / too many ties: histogram passes go on
    const u64 pre0 = s_pre[0];
    const double *p = a + row * n;
    u64 *cand = ws.cand + (size_t)row * kQCap;
    const int hshift = q_shift(1);  // the prefix is the top 22 bits
    const i64 per_cta = (n + gridDim.x - 1) / gridDim.x;
    const i64 lo = (i64)blockIdx.x * per_cta;
    const i64 hi = lo + per_cta < n ? lo + per_cta : n;
    constexpr int U = 8;
    for (i64 base = lo; base < hi; base += (i64)U * kQThreads) {
        double xs[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 i = base + (i64)u * kQThreads + tid;
            xs[u] = i < hi ? __ldcs(p + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool in = base + (i64)u * kQThreads + tid < hi;
            const u64 key = quant_key(xs[u]);
            const u64 head = key >> hshift;
            bool hit = in && head == pre0;
            for (int c = 1; c < nd; c++) hit = hit || (in && head == s_pre[c]);
            if (hit) {
                const unsigned pos = atomicAdd(&ws.ncand[row], 1u);
                if (pos < (unsigned)kQCap) cand[pos] = key;
            }
        }
    }
}

// One CTA per row: sort the candidate list, read the order statistics off.  The candidates of
// different prefixes occupy disjoint key ranges, so target t is at (first key >= prefix << 42) + rank
// inside its bucket.
__global__ void __launch_bounds__(kQThreads) quant_candidates_kernel(QuantWs ws, int T2) {
    __shared__ u64 keys[kQCap];
    __shared__ int s_tab[2 * kQMaxQ];
    __shared__ u64 s_pre[2 * kQMaxQ];
    __shared__ unsigned s_cnt[2 * kQMaxQ];
    const int tid = threadIdx.x;
    const i64 row = blockIdx.x;
    const int nd = quant_distinct(ws, row, T2, s_pre, s_tab, s_cnt, nullptr);
    unsigned total = 0;
    for (int c = 0; c < nd; c++) total += s_cnt[c];
    if (nd == 0 || total > (unsigned)kQCap) return;
    const int cnt = (int)min(ws.ncand[row], (unsigned)kQCap);  // == total
    int npad = 2;
    while (npad < cnt) npad <<= 1;
    const u64 *cand = ws.cand + (size_t)row * kQCap;
    for (int i = tid; i < npad; i += kQThreads) keys[i] = i < cnt ? cand[i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += kQThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int pp = i | j;
                const bool up = (i & k) == 0;
                const u64 x = keys[i], y = keys[pp];
                if ((x > y) == up) {
                    keys[i] = y;
                    keys[pp] = x;
                }
            }
            __syncthreads();
        }
    }
    if (tid < T2) {
        const i64 rem = ws.remaining[row * T2 + tid];
        if (rem >= 0) {
            const u64 first = ws.prefix[row * T2 + tid] << q_shift(1);
            int lo = 0, hi = cnt;  // first index with keys[idx] >= first
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (keys[mid] < first) lo = mid + 1; else hi = mid;
            }
            const int at = lo + (int)rem;
            ws.prefix[row * T2 + tid] = keys[at < cnt ? at : cnt - 1];
        }
    }
    if (tid == 0) ws.done[row] = 1;
}

__global__ void quant_finish_kernel(QuantWs ws, const double *__restrict__ q, double *__restrict__ out, i64 rows, int m,
                                    int m_total, int t_off) {
    const i64 gid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= rows * m) return;
    const i64 row = gid / m;
    const int t = (int)(gid % m);
    const i64 valid = ws.valid[row];
    const double qq = q[t];
    double r = quiet_nan<double>();
    if (valid > 0 && qq == qq) {
        double rank;
        i64 lo, hi;
        quant_rank(valid, qq, rank, lo, hi);
        const double fv = key_to_double(ws.prefix[(row * m + t) * 2]);
        const double cv = key_to_double(ws.prefix[(row * m + t) * 2 + 1]);
        r = quant_interpolate(fv, cv, rank, lo);
    }
    out[row * m_total + t_off + t] = r;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct QuantLayout {
    size_t valid, prefix, remaining, bucket, alias, done, ncand, hist, cand, zero_bytes, total;
};
QuantLayout quant_layout(i64 rows, i64 m) {
    QuantLayout l;
    const size_t T2 = 2 * (size_t)(m < kQChunk ? m : kQChunk);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += align256(bytes);
        return at;
    };
    l.valid = take((size_t)rows * 8);
    l.prefix = take((size_t)rows * T2 * 8);
    l.remaining = take((size_t)rows * T2 * 8);
    l.bucket = take((size_t)rows * T2 * 4);
    l.alias = take((size_t)rows * T2 * 4);
    l.done = take((size_t)rows * 4);
    l.ncand = take((size_t)rows * 4);
    l.hist = take((size_t)rows * T2 * kQBins * 4);
    l.zero_bytes = o;  // everything up to here starts at zero
    l.cand = take((size_t)rows * kQCap * 8);
    l.total = o;
    return l;
}

}  // namespace
}  // namespace nbg

using namespace nbg;

extern "C" size_t nbg_quantile_workspace_bytes(int64_t rows, int64_t n, int64_t m) {
    if (rows <= 0 || n <= kQSortMax || m <= 0) return 0;
    return quant_layout(rows, m).total + 256;
}

extern "C" int nbg_quantile(const void *a, const void *q, void *out, int64_t rows, int64_t n, int64_t m,
                            void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || n < 0 || m < 0) return fail(NBG_ERR_BAD_ARG, "nbg_quantile: negative extent");
    if (m > kQMaxQ) return fail(NBG_ERR_BAD_ARG, "nbg_quantile: at most 16 quantiles per call");
    if (rows == 0 || m == 0) return NBG_OK;
    if (out == nullptr || q == nullptr || (a == nullptr && n > 0)) return fail(NBG_ERR_BAD_ARG, "nbg_quantile: null pointer");
    if (rows > 0x7fffffff || n >= ((int64_t)1 << 31))
        return fail(NBG_ERR_BAD_ARG, "nbg_quantile: rows and n must be below 2^31 (the reference indexes with int32)");
    const double *ad = (const double *)a, *qd = (const double *)q;
    if (n <= 1024 && !getenv("NBG_QUANT_NOWARP")) {
        const unsigned blocks = (unsigned)((rows + 3) / 4);
        if (n <= 32)
            quant_warp_sort_kernel<1><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        else if (n <= 64)
            quant_warp_sort_kernel<2><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        else if (n <= 128)
            quant_warp_sort_kernel<4><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        else if (n <= 256)
            quant_warp_sort_kernel<8><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        else if (n <= 512)
            quant_warp_sort_kernel<16><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        else
            quant_warp_sort_kernel<32><<<blocks, 128, 0, stream>>>(ad, qd, (double *)out, rows, n, (int)m);
        return check_launch("nbg_quantile warp sort");
    }
    if (n <= kQSortMax && n > 512) {
        const size_t smem = (((size_t)n * 8 + 15) & ~(size_t)15) + (size_t)2 * m * 256 * sizeof(unsigned);
        int rc = allow_big_smem(quant_smem_select_kernel, "nbg_quantile: cudaFuncSetAttribute");
        if (rc) return rc;
        quant_smem_select_kernel<<<(unsigned)rows, n <= 2048 ? 128 : kQThreads, smem, stream>>>(ad, qd, (double *)out, n,
                                                                                          (int)m);
        return check_launch("nbg_quantile smem select");
    }
    if (n <= kQSortMax) {
        int npad = 2;
        while (npad < n) npad <<= 1;
        int threads = npad / 2 < 32 ? 32 : (npad / 2 > kQThreads ? kQThreads : npad / 2);
        quant_sort_kernel<<<(unsigned)rows, threads, (size_t)npad * 8, stream>>>(ad, qd, (double *)out, n, npad, (int)m);
        return check_launch("nbg_quantile sort");
    }
    const QuantLayout l = quant_layout(rows, m);
    if (workspace == nullptr || workspace_bytes < l.total + 256)
        return fail(NBG_ERR_WORKSPACE, "nbg_quantile: workspace smaller than nbg_quantile_workspace_bytes()");
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    QuantWs ws{(i64 *)(base + l.valid), (u64 *)(base + l.prefix), (i64 *)(base + l.remaining), (unsigned *)(base + l.bucket),
               (int *)(base + l.alias), (unsigned *)(base + l.hist), (int *)(base + l.done), (unsigned *)(base + l.ncand),
               (u64 *)(base + l.cand)};
    if (rows > 65535) return fail(NBG_ERR_UNSUPPORTED, "nbg_quantile: more than 65535 rows longer than 4096 elements");
    // about one wave of CTAs over all rows; every CTA at least 4096 elements
    int64_t segs = ((int64_t)kNumSMs * 8 + rows - 1) / rows;
    const int64_t max_segs = (n + 4095) / 4096;
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    int rc = allow_big_smem(quant_hist_kernel, "nbg_quantile: cudaFuncSetAttribute");
    if (rc) return rc;
    for (int64_t q0 = 0; q0 < m; q0 += kQChunk) {  // rounds of <= 12 quantiles: 24 histograms of 8 KB in shared memory
        const int mc = (int)(m - q0 < kQChunk ? m - q0 : kQChunk);
        const int T2 = 2 * mc;
        const size_t hist_bytes = (size_t)rows * T2 * kQBins * sizeof(unsigned);
        if ((rc = check_cuda(cudaMemsetAsync(base, 0, l.zero_bytes, stream), "nbg_quantile: workspace memset"))) return rc;
        const int64_t sel_threads = rows * T2 * 32;
        for (int pass = 0; pass < kQPasses; pass++) {
            if (pass == 2) {
                // the third read compacts the candidates; rows that fit are finished from their list
                quant_compact_kernel<<<dim3((unsigned)segs, (unsigned)rows), kQThreads, 0, stream>>>(ad, ws, n, T2);
                if ((rc = check_launch("nbg_quantile compact"))) return rc;
                quant_candidates_kernel<<<(unsigned)rows, kQThreads, 0, stream>>>(ws, T2);
                if ((rc = check_launch("nbg_quantile candidates"))) return rc;
            }
            if (pass > 0 && (rc = check_cuda(cudaMemsetAsync(ws.hist, 0, hist_bytes, stream), "nbg_quantile: histogram memset")))
                return rc;
            const size_t smem = (size_t)(pass == 0 ? 1 : T2) * kQBins * sizeof(unsigned);
            quant_hist_kernel<<<dim3((unsigned)segs, (unsigned)rows), kQThreads, smem, stream>>>(ad, ws, n, T2, pass);
            if ((rc = check_launch("nbg_quantile hist"))) return rc;
            quant_select_kernel<<<(unsigned)((sel_threads + 127) / 128), 128, 0, stream>>>(ws, qd + q0, rows, T2, pass);
            if ((rc = check_launch("nbg_quantile select"))) return rc;
        }
        quant_finish_kernel<<<(unsigned)((rows * mc + 127) / 128), 128, 0, stream>>>(ws, qd + q0, (double *)out, rows, mc, (int)m,
                                                                                    (int)q0);
        if ((rc = check_launch("nbg_quantile finish"))) return rc;
    }
    return NBG_OK;
}
