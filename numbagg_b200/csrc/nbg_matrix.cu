// nbg_matrix.cu -- pairwise-complete covariance / correlation matrices: static
// (nancorrmatrix, nancovmatrix: numbagg/funcs.py:338-532), moving-window (move_corrmatrix,
// move_covmatrix: numbagg/moving_matrix.py:16-204) and exponentially weighted
// (move_exp_nancorrmatrix, move_exp_nancovmatrix: moving_matrix.py:207-432).
//
// First version: one thread per (batch item, i, j) pair runs the reference's loop body for
// that pair, in the reference's order and in the reference's types -- the running sums are
// kept in the INPUT dtype (numbagg allocates them with dtype=a.dtype), products of two inputs
// are rounded to it, everything divided by an int64 count is double -- so results are
// bit-identical to numbagg for float32 and float64 alike.  The pairs of one (batch, i) row
// sit in consecutive threads: loads of the j-th variable and stores of out[..., i, j] are
// coalesced.  Parallelism is batch x n_vars^2 threads only (no split along the observation
// axis yet: that changes the summation order and gives up bit-exactness for float32).
#include "nbg_common.cuh"

namespace nbg {
namespace {

typedef long long i64;

constexpr int kMatThreads = 128;

template <typename T>
__device__ __forceinline__ T mul_t(T a, T b);
template <>
__device__ __forceinline__ float mul_t<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_t<double>(double a, double b) { return __dmul_rn(a, b); }

// a: (batch, nv, no); out: (batch, nv, nv).  funcs.py:423-474 / 503-532.
template <typename T, bool CORR>
__global__ void __launch_bounds__(kMatThreads) mat_static_kernel(const T *__restrict__ a, T *__restrict__ out, i64 batch,
                                                                 int nv, i64 no) {
    const i64 gid = (i64)blockIdx.x * kMatThreads + threadIdx.x;
    const i64 q = (i64)nv * nv;
    if (gid >= batch * q) return;
    const i64 bi = gid / q;
    const int p = (int)(gid % q), i = p / nv, j = p % nv;
    if (j < i) return;  // upper triangle, mirrored below (funcs.py:436)
    const T *ri = a + (bi * nv + i) * no, *rj = a + (bi * nv + j) * no;
    T si = 0, sj = 0, sqi = 0, sqj = 0, sij = 0;
    i64 count = 0;
    for (i64 k = 0; k < no; k++) {
        const T vi = ri[k], vj = rj[k];
        if (is_nan(vi) || is_nan(vj)) continue;
        si += vi;
        sj += vj;
        if (CORR) {
            sqi += mul_t(vi, vi);
            sqj += mul_t(vj, vj);
        }
        sij += mul_t(vi, vj);
        count += 1;
    }
    T res = quiet_nan<T>();
    if (count > 1) {
        const double c = (double)count, c1 = (double)(count - 1);
        const double mi = (double)si / c, mj = (double)sj / c;
        const double cov = dsub((double)sij / c, dmul(mi, mj));
        const double cov_u = dmul(cov, c) / c1;
        if (CORR) {
            const double vi = dsub((double)sqi / c, dmul(mi, mi));
            const double vj = dsub((double)sqj / c, dmul(mj, mj));
            const double vi_u = dmul(vi, c) / c1, vj_u = dmul(vj, c) / c1;
            if (vi_u > 0 && vj_u > 0) res = (T)(cov_u / sqrt(dmul(vi_u, vj_u)));
        } else {
            res = (T)cov_u;
        }
    }
    out[bi * q + (i64)i * nv + j] = res;
    out[bi * q + (i64)j * nv + i] = res;
}

// a: (batch, no, nv); out: (batch, no, nv, nv).  moving_matrix.py:52-124 / 143-204.
template <typename T, bool CORR>
__global__ void __launch_bounds__(kMatThreads) mat_move_kernel(const T *__restrict__ a, T *__restrict__ out, i64 batch,
                                                               i64 no, int nv, i64 window, i64 min_count) {
    const i64 gid = (i64)blockIdx.x * kMatThreads + threadIdx.x;
    const i64 q = (i64)nv * nv;
    if (gid >= batch * q) return;
    const i64 bi = gid / q;
    const int p = (int)(gid % q), i = p / nv, j = p % nv;
    const T *ab = a + bi * no * nv;
    T *ob = out + bi * no * q + p;
    if (min_count < 1) min_count = 1;
    const i64 corr_min = min_count > 2 ? min_count : 2;
    T si = 0, sj = 0, sqi = 0, sqj = 0, pr = 0;
    i64 n = 0;
    for (i64 t = 0; t < no; t++) {
        if (t >= window) {
            const T vi = ab[(t - window) * nv + i], vj = ab[(t - window) * nv + j];
            if (!(is_nan(vi) || is_nan(vj))) {
                si -= vi;
                sj -= vj;
                if (CORR) {
                    sqi -= mul_t(vi, vi);
                    sqj -= mul_t(vj, vj);
                }
                pr -= mul_t(vi, vj);
                n -= 1;
            }
        }
        {
            const T vi = ab[t * nv + i], vj = ab[t * nv + j];
            if (!(is_nan(vi) || is_nan(vj))) {
                si += vi;
                sj += vj;
                if (CORR) {
                    sqi += mul_t(vi, vi);
                    sqj += mul_t(vj, vj);
                }
                pr += mul_t(vi, vj);
                n += 1;
            }
        }
        T res = quiet_nan<T>();
        const double c = (double)n;
        if (CORR) {
            if (n >= corr_min) {
                const double mi = (double)si / c, mj = (double)sj / c;
                const double vi = dsub((double)sqi / c, dmul(mi, mi));
                const double vj = dsub((double)sqj / c, dmul(mj, mj));
                const double cov = dsub((double)pr / c, dmul(mi, mj));
                if (vi > 0 && vj > 0) res = (T)(cov / sqrt(dmul(vi, vj)));
            }
        } else if (n >= min_count && n > 1) {
            const double mi = (double)si / c, mj = (double)sj / c;
            res = (T)(dmul(dsub((double)pr / c, dmul(mi, mj)), c) / (double)(n - 1));
        }
        ob[t * q] = res;
    }
}

// ---- static forms, long observation axes: partial sums per segment, folded in order ------------------
// Thread per (batch item, segment, pair with j >= i): the reference's accumulation over the segment's
// observations, in the reference's types; a second kernel folds the segments' partial sums left to
// right (still in the input dtype) and applies the reference's read-out.  Like every segmented form
// here this agrees with the reference to the rounding of its running sums; axes < 8192 keep the
// single-pass kernel (bit-identical).  part: [batch][nseg][6][npairs] doubles (float sums are exact there).
__device__ __forceinline__ void tri_pair_static(int k, int nv, int &i, int &j) {
    int row = 0, rem = k;
    while (rem >= nv - row) {
        rem -= nv - row;
        row++;
    }
    i = row;
    j = row + rem;
}

template <typename T, bool CORR>
__global__ void __launch_bounds__(256) mat_static_part_kernel(const T *__restrict__ a, double *__restrict__ part, i64 batch,
                                                              int nv, i64 no, i64 seg_len, int nseg, int blocks_per_seg) {
    const i64 np = (i64)nv * (nv + 1) / 2;
    const int pb = blockIdx.x % blocks_per_seg;
    const i64 sb = blockIdx.x / blocks_per_seg;
    const int seg = (int)(sb % nseg);
    const i64 bi = sb / nseg;
    const int k = pb * 256 + threadIdx.x;
    if (k >= np || bi >= batch) return;
    int i, j;
    tri_pair_static(k, nv, i, j);
    const i64 k0 = (i64)seg * seg_len, k1 = k0 + seg_len < no ? k0 + seg_len : no;
    const T *ri = a + (bi * nv + i) * no, *rj = a + (bi * nv + j) * no;
    T si = 0, sj = 0, sqi = 0, sqj = 0, sij = 0;
    i64 count = 0;
    for (i64 t = k0; t < k1; t++) {
        const T vi = ri[t], vj = rj[t];
        if (is_nan(vi) || is_nan(vj)) continue;
        si += vi;
        sj += vj;
        if (CORR) {
            sqi += mul_t(vi, vi);
            sqj += mul_t(vj, vj);
        }
        sij += mul_t(vi, vj);
        count += 1;
    }
    double *o = part + ((bi * nseg + seg) * 6) * np + k;
    o[0 * np] = (double)si, o[1 * np] = (double)sj, o[2 * np] = (double)sqi, o[3 * np] = (double)sqj;
    o[4 * np] = (double)sij, o[5 * np] = (double)count;
}

template <typename T, bool CORR>
__global__ void __launch_bounds__(256) mat_static_fold_kernel(const double *__restrict__ part, T *__restrict__ out, i64 batch,
                                                              int nv, int nseg) {
    const i64 np = (i64)nv * (nv + 1) / 2;
    const i64 gid = (i64)blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * np) return;
    const i64 bi = gid / np;
    const int k = (int)(gid % np);
    int i, j;
    tri_pair_static(k, nv, i, j);
    T si = 0, sj = 0, sqi = 0, sqj = 0, sij = 0;
    i64 count = 0;
    for (int seg = 0; seg < nseg; seg++) {
        const double *o = part + ((bi * nseg + seg) * 6) * np + k;
        si += (T)o[0 * np];
        sj += (T)o[1 * np];
        if (CORR) {
            sqi += (T)o[2 * np];
            sqj += (T)o[3 * np];
        }
        sij += (T)o[4 * np];
        count += (i64)o[5 * np];
    }
    T res = quiet_nan<T>();
    if (count > 1) {  // funcs.py:423-474 / 503-532, as in mat_static_kernel
        const double c = (double)count, c1 = (double)(count - 1);
        const double mi = (double)si / c, mj = (double)sj / c;
        const double cov = dsub((double)sij / c, dmul(mi, mj));
        const double cov_u = dmul(cov, c) / c1;
        if (CORR) {
            const double vi = dsub((double)sqi / c, dmul(mi, mi));
            const double vj = dsub((double)sqj / c, dmul(mj, mj));
            const double vi_u = dmul(vi, c) / c1, vj_u = dmul(vj, c) / c1;
            if (vi_u > 0 && vj_u > 0) res = (T)(cov_u / sqrt(dmul(vi_u, vj_u)));
        } else {
            res = (T)cov_u;
        }
    }
    const i64 q = (i64)nv * nv;
    out[bi * q + (i64)i * nv + j] = res;
    out[bi * q + (i64)j * nv + i] = res;
}

// ---- long observation axes: segments -------------------------------------------------------------
// One thread per (batch item, segment, i, j).  A segment rebuilds the state of its pair from the
// `window` observations before it (additions only, in order) and then runs the reference's
// recurrence.  The reference never re-syncs its running sums, so the results agree to rounding of
// the running sums, not bit for bit; short axes keep the single-segment kernel
// above.  Quotients by the integer count come from a per-CTA table of correctly rounded reciprocals
// plus the residual correction q' = q + (a - c q) y (Markstein: correctly rounded a / c), i.e. they
// round exactly like the reference's divisions at a quarter of the FP64 instructions.
constexpr int kMatSegThreads = 256;
constexpr int kMatRcpMax = 4096;  // windows up to this long use the reciprocal table

// The matrices are symmetric and so is every loop body above (the reference computes all nv^2 pairs
// separately; swapping i and j only swaps commutative operands), so the segmented kernels compute
// the nv (nv + 1) / 2 pairs with j >= i and store each result twice.  k -> (i, j), row-major over
// the upper triangle: consecutive threads are consecutive j of one row (coalesced primary store).
__device__ __forceinline__ void tri_pair(int k, int nv, int &i, int &j) {
    int row = 0, rem = k;
    while (rem >= nv - row) {
        rem -= nv - row;
        row++;
    }
    i = row;
    j = row + rem;
}

// (tried: prefetch.global.L1 of the rows 8 steps ahead -- slower, 1.10 vs 0.81 ms on 200 000 x 32 float64)
template <typename T, bool CORR, bool TABLE, bool STAGE>
__global__ void __launch_bounds__(kMatSegThreads) mat_move_seg_kernel(const T *__restrict__ a, T *__restrict__ out, i64 batch,
                                                                      i64 no, int nv, i64 window, i64 min_count, i64 seg_len,
                                                                      int nseg, int blocks_per_seg, int OB) {
    extern __shared__ double rc_tab[];  // rc_tab[c] = 1 / c
    if (TABLE) {
        for (int c = threadIdx.x; c <= (int)window; c += kMatSegThreads) rc_tab[c] = c ? 1.0 / (double)c : 0.0;
        __syncthreads();
    }
    const i64 q = (i64)nv * nv;
    const int npairs = nv * (nv + 1) / 2;
    const int pb = blockIdx.x % blocks_per_seg;
    const i64 sb = blockIdx.x / blocks_per_seg;
    const int seg = (int)(sb % nseg);
    const i64 bi = sb / nseg;
    const int k0 = pb * kMatSegThreads + threadIdx.x;
    const bool active = k0 < npairs;  // (bi < batch by construction of the grid)
    if (!STAGE && !active) return;
    const int k = active ? k0 : 0;    // idle threads of the staged form still load and synchronise
    int i, j;
    tri_pair(k, nv, i, j);
    const T *ab = a + bi * no * nv;
    T *ob = out + bi * no * q;
    const int oij = i * nv + j, oji = j * nv + i;
    if (min_count < 1) min_count = 1;
    const int corr_min = (int)(min_count > 2 ? min_count : 2);
    const int cov_min = (int)(min_count > 2 ? min_count : 2);  // n >= min_count && n > 1
    const i64 t0 = (i64)seg * seg_len;
    const i64 t1 = t0 + seg_len < no ? t0 + seg_len : no;
    T si = 0, sj = 0, sqi = 0, sqj = 0, pr = 0;
    int n = 0;
    for (i64 t = t0 > window ? t0 - window : 0; t < t0; t++) {
        const T vi = ab[t * nv + i], vj = ab[t * nv + j];
        if (!(is_nan(vi) || is_nan(vj))) {
            si += vi;
            sj += vj;
            if (CORR) {
                sqi += mul_t(vi, vi);
                sqj += mul_t(vj, vj);
            }
            pr += mul_t(vi, vj);
            n += 1;
        }
    }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    // one step of the reference's loop body: leaving pair (NaN before the window is full), entering
    // pair, read-out, both stores
    auto do_step = [&](T ti, T tj, T li, T lj, T *o) {
        {
            const T vi = ti, vj = tj;
            if (!(is_nan(vi) || is_nan(vj))) {
                si -= vi;
                sj -= vj;
                if (CORR) {
                    sqi -= mul_t(vi, vi);
                    sqj -= mul_t(vj, vj);
                }
                pr -= mul_t(vi, vj);
                n -= 1;
            }
        }
        {
            const T vi = li, vj = lj;
            if (!(is_nan(vi) || is_nan(vj))) {
                si += vi;
                sj += vj;
                if (CORR) {
                    sqi += mul_t(vi, vi);
                    sqj += mul_t(vj, vj);
                }
                pr += mul_t(vi, vj);
                n += 1;
            }
        }
        T res = quiet_nan<T>();
        if (TABLE) {
            // quotients by the count: q' = q + (a - c q) y with y = the correctly rounded 1 / c rounds like
            // the division itself (Markstein) unless a quotient is tiny or not finite -- ONE test for all
            // of them, then the IEEE path
            if (n >= (CORR ? corr_min : cov_min)) {
                const double c = (double)n, y = rc_tab[n];
                auto qd = [&](double x) {
                    const double q0 = x * y;
                    return fma(fma(-c, q0, x), y, q0);
                };
                const double mi = qd((double)si), mj = qd((double)sj), mp = qd((double)pr);
                double small = fmin(fmin(fabs(mi), fabs(mj)), fabs(mp)), big = fabs(mi) + fabs(mj) + fabs(mp);
                if (CORR) {
                    const double qi = qd((double)sqi), qj = qd((double)sqj);
                    small = fmin(small, fmin(fabs(qi), fabs(qj)));
                    big += fabs(qi) + fabs(qj);
                    if (small > 1e-290 && big < kInf) {
                        const double vi = dsub(qi, dmul(mi, mi)), vj = dsub(qj, dmul(mj, mj));
                        const double cov = dsub(mp, dmul(mi, mj));
                        if (vi > 0 && vj > 0) res = (T)(cov / sqrt(dmul(vi, vj)));
                    } else {
                        const double mi2 = (double)si / c, mj2 = (double)sj / c;
                        const double vi = dsub((double)sqi / c, dmul(mi2, mi2));
                        const double vj = dsub((double)sqj / c, dmul(mj2, mj2));
                        const double cov = dsub((double)pr / c, dmul(mi2, mj2));
                        if (vi > 0 && vj > 0) res = (T)(cov / sqrt(dmul(vi, vj)));
                    }
                } else {
                    const double c1 = (double)(n - 1), y1 = rc_tab[n - 1];
                    const double num = dmul(dsub(mp, dmul(mi, mj)), c);
                    const double q0 = num * y1;
                    const double r = fma(fma(-c1, q0, num), y1, q0);
                    if (small > 1e-290 && big < kInf && fabs(q0) > 1e-290)
                        res = (T)r;
                    else
                        res = (T)(dmul(dsub((double)pr / c, dmul((double)si / c, (double)sj / c)), c) / c1);
                }
            }
        } else {
            const double c = (double)n;
            if (CORR) {
                if (n >= corr_min) {
                    const double mi = (double)si / c, mj = (double)sj / c;
                    const double vi = dsub((double)sqi / c, dmul(mi, mi));
                    const double vj = dsub((double)sqj / c, dmul(mj, mj));
                    const double cov = dsub((double)pr / c, dmul(mi, mj));
                    if (vi > 0 && vj > 0) res = (T)(cov / sqrt(dmul(vi, vj)));
                }
            } else if (n >= cov_min) {
                const double mi = (double)si / c, mj = (double)sj / c;
                res = (T)(dmul(dsub((double)pr / c, dmul(mi, mj)), c) / (double)(n - 1));
            }
        }
        __stcs(o + oij, res);
        if (i != j) __stcs(o + oji, res);
    };
    const T kNaN = quiet_nan<T>();
    T *o = ob + t0 * q;
    if constexpr (STAGE) {
        // The observation rows of a chunk of OB steps -- entering rows [tc, tc + OB) and leaving rows
        // [tc - window, ...) -- are staged in shared memory by the whole CTA (all of its threads work on
        // the same segment), double-buffered: the next chunk travels global -> registers while this one
        // is consumed, registers -> shared memory at its end.  Without it every step starts with two
        // L2 round trips (half of the stall samples of the unstaged form).
        T *stg = reinterpret_cast<T *>(rc_tab + (TABLE ? (((int)window + 2) & ~1) : 0));
        const int ce = OB * nv;  // elements per buffer and stream, <= 4 * kMatSegThreads
        const i64 total = no * nv;
        auto fetch = [&](i64 tc, int which, int e) -> T {
            const i64 g = (which ? tc - window : tc) * nv + e;
            return (g >= 0 && g < total) ? ab[g] : kNaN;
        };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = threadIdx.x + u * kMatSegThreads;
            if (e < ce) {
                stg[e] = fetch(t0, 0, e);
                stg[ce + e] = fetch(t0, 1, e);
            }
        }
        __syncthreads();
        int b = 0;
        for (i64 tc = t0; tc < t1; tc += OB, b ^= 1) {
            T pl[4], pt[4];
            const bool more = tc + OB < t1;
            if (more) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = threadIdx.x + u * kMatSegThreads;
                    pl[u] = e < ce ? fetch(tc + OB, 0, e) : kNaN;
                    pt[u] = e < ce ? fetch(tc + OB, 1, e) : kNaN;
                }
            }
            if (active) {
                const T *L = stg + (size_t)(2 * b) * ce, *R = L + ce;
                const int ns = (int)(t1 - tc < OB ? t1 - tc : OB);
                for (int st = 0; st < ns; st++, o += q) do_step(R[st * nv + i], R[st * nv + j], L[st * nv + i], L[st * nv + j], o);
            }
            if (more) {
                T *W = stg + (size_t)(2 * (b ^ 1)) * ce;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = threadIdx.x + u * kMatSegThreads;
                    if (e < ce) {
                        W[e] = pl[u];
                        W[ce + e] = pt[u];
                    }
                }
            }
            __syncthreads();
        }
    } else {
        const T *lead = ab + t0 * nv, *trail = ab + (t0 - window) * nv;  // trail is only dereferenced once t >= window
        // the observations of step t + 1 are fetched while step t computes
        T n_li = t0 < t1 ? lead[i] : kNaN, n_lj = t0 < t1 ? lead[j] : kNaN;
        T n_ti = t0 >= window && t0 < t1 ? trail[i] : kNaN, n_tj = t0 >= window && t0 < t1 ? trail[j] : kNaN;
        for (i64 t = t0; t < t1; t++, lead += nv, trail += nv, o += q) {
            const T li = n_li, lj = n_lj, ti = n_ti, tj = n_tj;
            if (t + 1 < t1) {
                n_li = lead[nv + i], n_lj = lead[nv + j];
                if (t + 1 >= window) n_ti = trail[nv + i], n_tj = trail[nv + j];
            }
            do_step(ti, tj, li, lj, o);
        }
    }
}

// ---- long observation axes, exponential weights: segments with carried state ----------------------
// The eight running sums of a pair obey s <- decay_t * s + contribution_t (psw2: decay_t^2), so a
// segment is the affine map s -> D s + U with D = prod decay_t (the same for every pair) and
// U = the segment run from zero.  Pass A runs every (segment, pair) from zero and stores U and D;
// `mat_exp_carry_kernel` folds them left to right into the state each segment starts from; pass B
// runs the reference's loop body from that state and writes the outputs.  Like the segmented
// windows above this agrees with the reference to the rounding of the running sums (which the
// reference keeps in the input dtype), not bit for bit; short axes keep the single-segment kernel.
constexpr int kMatExpStates = 8;  // si sj sqi sqj pr pw psw psw2

template <typename T, bool CORR>
struct MatExpState {
    T si = 0, sj = 0, sqi = 0, sqj = 0, pr = 0, pw = 0, psw = 0, psw2 = 0;
    __device__ __forceinline__ void step(T alpha_t, T vi, T vj) {
        const double decay = dsub(1.0, (double)alpha_t);
        const double decay2 = dmul(decay, decay);
        si = (T)dmul((double)si, decay);
        sj = (T)dmul((double)sj, decay);
        if (CORR) {
            sqi = (T)dmul((double)sqi, decay);
            sqj = (T)dmul((double)sqj, decay);
        }
        pr = (T)dmul((double)pr, decay);
        pw = (T)dmul((double)pw, decay);
        psw = (T)dmul((double)psw, decay);
        psw2 = (T)dmul((double)psw2, decay2);
        if (!(is_nan(vi) || is_nan(vj))) {
            si += vi;
            sj += vj;
            if (CORR) {
                sqi += mul_t(vi, vi);
                sqj += mul_t(vj, vj);
            }
            pr += mul_t(vi, vj);
            pw += alpha_t;
            psw = (T)dadd((double)psw, 1.0);
            psw2 = (T)dadd((double)psw2, 1.0);
        }
    }
    __device__ __forceinline__ T read_out(T min_weight) const {
        if constexpr (sizeof(T) == 8) {
            // double: every quotient of the reference's read-out through one reciprocal per distinct
            // divisor + the residual correction (correctly rounded, nbg_common.cuh) -- 3 instructions per
            // quotient instead of an IEEE division each
            const double n = psw, n2 = dmul(psw, psw);
            if (rcp_ok(n) && rcp_ok(n2)) {
                T res = quiet_nan<T>();
                if (!(psw > 0)) return res;  // bias stays 0
                const double bias = dsub(1.0, qdiv(psw2, n2, fast_rcp(n2)));
                if (pw >= min_weight && bias > 0) {
                    const double y = fast_rcp(n);
                    const double mi = qdiv(si, n, y), mj = qdiv(sj, n, y);
                    const double cov_b = dsub(qdiv(pr, n, y), dmul(mi, mj));
                    if (rcp_ok(bias)) {
                        const double yb = fast_rcp(bias);
                        if (CORR) {
                            const double vi_b = dsub(qdiv(sqi, n, y), dmul(mi, mi));
                            const double vj_b = dsub(qdiv(sqj, n, y), dmul(mj, mj));
                            const double vvi = qdiv(vi_b, bias, yb), vvj = qdiv(vj_b, bias, yb);
                            const double cov = qdiv(cov_b, bias, yb);
                            if (vvi > 0 && vvj > 0) res = (T)fdiv(cov, sqrt(dmul(vvi, vvj)));
                        } else {
                            res = (T)qdiv(cov_b, bias, yb);
                        }
                    } else if (CORR) {
                        const double vi_b = dsub(qdiv(sqi, n, y), dmul(mi, mi));
                        const double vj_b = dsub(qdiv(sqj, n, y), dmul(mj, mj));
                        const double vvi = vi_b / bias, vvj = vj_b / bias, cov = cov_b / bias;
                        if (vvi > 0 && vvj > 0) res = (T)(cov / sqrt(dmul(vvi, vvj)));
                    } else {
                        res = (T)(cov_b / bias);
                    }
                }
                return res;
            }
        }
        double bias = 0.0;
        if (psw > (T)0) bias = dsub(1.0, (double)(T)(psw2 / mul_t(psw, psw)));
        T res = quiet_nan<T>();
        if (pw >= min_weight && bias > 0) {
            const T n = psw;
            const T mi = si / n, mj = sj / n;
            const T cov_b = (T)(pr / n) - mul_t(mi, mj);
            if (CORR) {
                const T vi_b = (T)(sqi / n) - mul_t(mi, mi);
                const T vj_b = (T)(sqj / n) - mul_t(mj, mj);
                const double vvi = (double)vi_b / bias, vvj = (double)vj_b / bias;
                const double cov = (double)cov_b / bias;
                if (vvi > 0 && vvj > 0) res = (T)(cov / sqrt(dmul(vvi, vvj)));
            } else {
                res = (T)((double)cov_b / bias);
            }
        }
        return res;
    }
};

// carry: [batch][nseg][kMatExpStates][npairs] doubles; decays: [batch][nseg][2] (D, D2)
template <typename T, bool CORR, bool PASS_A>
__global__ void __launch_bounds__(kMatSegThreads) mat_exp_seg_kernel(const T *__restrict__ a, const T *__restrict__ alpha,
                                                                     int alpha_per_item, T min_weight, T *__restrict__ out,
                                                                     i64 batch, i64 no, int nv, i64 seg_len, int nseg,
                                                                     int blocks_per_seg, double *__restrict__ carry,
                                                                     double *__restrict__ decays) {
    const i64 q = (i64)nv * nv;
    const i64 np = (i64)nv * (nv + 1) / 2;
    const int pb = blockIdx.x % blocks_per_seg;
    const i64 sb = blockIdx.x / blocks_per_seg;
    const int seg = (int)(sb % nseg);
    const i64 bi = sb / nseg;
    const int k = pb * kMatSegThreads + threadIdx.x;
    if (k >= np || bi >= batch) return;
    if (PASS_A && seg == nseg - 1) return;  // nobody starts from the last segment's end
    int i, j;
    tri_pair(k, nv, i, j);
    const T *al = alpha + (alpha_per_item ? bi * no : 0);
    const i64 t0 = (i64)seg * seg_len;
    const i64 t1 = t0 + seg_len < no ? t0 + seg_len : no;
    const T *lead = a + (bi * no + t0) * nv;
    double *cr = carry + ((bi * nseg + seg) * kMatExpStates) * np + k;
    MatExpState<T, CORR> st;
    if (!PASS_A && seg > 0) {
        st.si = (T)cr[0 * np], st.sj = (T)cr[1 * np], st.pr = (T)cr[4 * np];
        st.pw = (T)cr[5 * np], st.psw = (T)cr[6 * np], st.psw2 = (T)cr[7 * np];
        if (CORR) st.sqi = (T)cr[2 * np], st.sqj = (T)cr[3 * np];
    }
    if (PASS_A) {
        double D = 1.0, D2 = 1.0;
        T n_i = t0 < t1 ? lead[i] : (T)0, n_j = t0 < t1 ? lead[j] : (T)0, n_al = t0 < t1 ? al[t0] : (T)0;
        for (i64 t = t0; t < t1; t++, lead += nv) {
            const T alpha_t = n_al, vi = n_i, vj = n_j;
            if (t + 1 < t1) n_i = lead[nv + i], n_j = lead[nv + j], n_al = al[t + 1];
            st.step(alpha_t, vi, vj);
            if (k == 0) {
                const double decay = dsub(1.0, (double)alpha_t);
                D = dmul(D, decay);
                D2 = dmul(D2, dmul(decay, decay));
            }
        }
        cr[0 * np] = (double)st.si, cr[1 * np] = (double)st.sj, cr[4 * np] = (double)st.pr;
        cr[5 * np] = (double)st.pw, cr[6 * np] = (double)st.psw, cr[7 * np] = (double)st.psw2;
        if (CORR) cr[2 * np] = (double)st.sqi, cr[3 * np] = (double)st.sqj;
        if (k == 0) decays[(bi * nseg + seg) * 2] = D, decays[(bi * nseg + seg) * 2 + 1] = D2;
    } else {
        T *o = out + (bi * no + t0) * q;
        const int oij = i * nv + j, oji = j * nv + i;
        T n_i = t0 < t1 ? lead[i] : (T)0, n_j = t0 < t1 ? lead[j] : (T)0, n_al = t0 < t1 ? al[t0] : (T)0;
        for (i64 t = t0; t < t1; t++, lead += nv, o += q) {
            const T alpha_t = n_al, vi = n_i, vj = n_j;
            if (t + 1 < t1) n_i = lead[nv + i], n_j = lead[nv + j], n_al = al[t + 1];
            st.step(alpha_t, vi, vj);
            const T res = st.read_out(min_weight);
            __stcs(o + oij, res);
            if (i != j) __stcs(o + oji, res);
        }
    }
}

// U records -> incoming states, in place: thread per (batch item, state, pair)
__global__ void mat_exp_carry_kernel(double *__restrict__ carry, const double *__restrict__ decays, i64 batch, int nseg, i64 q) {
    const i64 gid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * kMatExpStates * q) return;
    const i64 p = gid % q;
    const int k = (int)((gid / q) % kMatExpStates);
    const i64 bi = gid / (q * kMatExpStates);
    double s = 0.0;
    const i64 stride = (i64)kMatExpStates * q;
    double *c0 = carry + ((bi * nseg) * kMatExpStates + k) * q + p;
    const double *dc = decays + bi * nseg * 2 + (k == 7 ? 1 : 0);
    constexpr int B = 8;  // records in flight: the loads of a batch are independent of the running state
    for (int seg0 = 0; seg0 < nseg; seg0 += B) {
        double u[B], d[B];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int seg = seg0 + b;
            u[b] = seg < nseg - 1 ? c0[seg * stride] : 0.0;
            d[b] = seg < nseg - 1 ? dc[seg * 2] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int seg = seg0 + b;
            if (seg < nseg) c0[seg * stride] = s;
            s = dadd(dmul(d[b], s), u[b]);
        }
    }
}

// a: (batch, no, nv); alpha: (no) or (batch, no); out: (batch, no, nv, nv).
// moving_matrix.py:247-331 / 372-432.
template <typename T, bool CORR>
__global__ void __launch_bounds__(kMatThreads) mat_exp_kernel(const T *__restrict__ a, const T *__restrict__ alpha,
                                                              int alpha_per_item, T min_weight, T *__restrict__ out,
                                                              i64 batch, i64 no, int nv) {
    const i64 gid = (i64)blockIdx.x * kMatThreads + threadIdx.x;
    const i64 q = (i64)nv * nv;
    if (gid >= batch * q) return;
    const i64 bi = gid / q;
    const int p = (int)(gid % q), i = p / nv, j = p % nv;
    const T *ab = a + bi * no * nv;
    const T *al = alpha + (alpha_per_item ? bi * no : 0);
    T *ob = out + bi * no * q + p;
    T si = 0, sj = 0, sqi = 0, sqj = 0, pr = 0, pw = 0, psw = 0, psw2 = 0;
    for (i64 t = 0; t < no; t++) {
        const T alpha_t = al[t];
        const double decay = dsub(1.0, (double)alpha_t);
        const double decay2 = dmul(decay, decay);
        // `arr *= decay` on arrays of the input dtype: through double, rounded back
        si = (T)dmul((double)si, decay);
        sj = (T)dmul((double)sj, decay);
        if (CORR) {
            sqi = (T)dmul((double)sqi, decay);
            sqj = (T)dmul((double)sqj, decay);
        }
        pr = (T)dmul((double)pr, decay);
        pw = (T)dmul((double)pw, decay);
        psw = (T)dmul((double)psw, decay);
        psw2 = (T)dmul((double)psw2, decay2);
        const T vi = ab[t * nv + i], vj = ab[t * nv + j];
        if (!(is_nan(vi) || is_nan(vj))) {
            si += vi;
            sj += vj;
            if (CORR) {
                sqi += mul_t(vi, vi);
                sqj += mul_t(vj, vj);
            }
            pr += mul_t(vi, vj);
            pw += alpha_t;
            psw = (T)dadd((double)psw, 1.0);
            psw2 = (T)dadd((double)psw2, 1.0);
        }
        double bias = 0.0;
        if (psw > (T)0) bias = dsub(1.0, (double)(T)(psw2 / mul_t(psw, psw)));
        T res = quiet_nan<T>();
        if (pw >= min_weight && bias > 0) {
            const T n = psw;
            const T mi = si / n, mj = sj / n;
            const T cov_b = (T)(pr / n) - mul_t(mi, mj);
            if (CORR) {
                const T vi_b = (T)(sqi / n) - mul_t(mi, mi);
                const T vj_b = (T)(sqj / n) - mul_t(mj, mj);
                const double vvi = (double)vi_b / bias, vvj = (double)vj_b / bias;
                const double cov = (double)cov_b / bias;
                if (vvi > 0 && vvj > 0) res = (T)(cov / sqrt(dmul(vvi, vvj)));
            } else {
                res = (T)((double)cov_b / bias);
            }
        }
        ob[t * q] = res;
    }
}

// Segments along the observation axis: only for long axes (short ones stay bit-identical to the
// reference).  All CTAs do the same amount of work, so the best grid is exactly ONE wave of resident
// CTAs (`resident` = SMs x occupancy of the kernel; 1.27 waves cost two): as many segments as fit,
// each at least 8 windows long so that rebuilding the window costs <= 1/8 extra.
constexpr i64 kMatSegMinObs = 8192;
struct MatSegs {
    i64 seg_len;
    int nseg;
};
static MatSegs mat_segments(i64 ctas_per_seg, i64 resident, i64 no, i64 window) {
    MatSegs s = {no, 1};
    if (getenv("NBG_MAT_NOSEG") || no < kMatSegMinObs || ctas_per_seg <= 0) return s;
    i64 nseg = resident / ctas_per_seg;
    if (nseg < 2) return s;
    i64 len = (no + nseg - 1) / nseg;
    const i64 min_len = window * 8 > 256 ? window * 8 : 256;
    if (len < min_len) len = min_len;
    if (len >= no) return s;
    s.seg_len = len;
    s.nseg = (int)((no + len - 1) / len);
    return s;
}
template <class K>
static i64 resident_ctas(K kern, int threads, size_t smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    return (i64)occ * kNumSMs;
}

// The carry workspace of the exponential forms comes from the device's stream-ordered pool.  By
// default the pool hands freed memory back to the driver at the next synchronisation, which turns
// every call after a sync into a real allocation (~1-2 ms measured); keep it cached instead.
static void keep_pool_memory() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = (uint64_t)1 << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

template <typename T>
int launch_matrix(int op, const void *a_, const void *alpha_, int alpha_per_item, double min_weight, void *out_, i64 batch,
                  i64 no, i64 nv, i64 window, i64 min_count, cudaStream_t stream) {
    const T *a = (const T *)a_;
    T *out = (T *)out_;
    const i64 threads = batch * nv * nv;
    if (threads <= 0) return NBG_OK;
    if ((threads + kMatThreads - 1) / kMatThreads > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: too many pairs");
    const unsigned grid = (unsigned)((threads + kMatThreads - 1) / kMatThreads);
    switch (op) {
        case NBG_MAT_NANCORR:
        case NBG_MAT_NANCOV: {
            const bool corr = op == NBG_MAT_NANCORR;
            const i64 np = nv * (nv + 1) / 2;
            const int bps = (int)((np + 255) / 256);
            const i64 resident = corr ? resident_ctas(mat_static_part_kernel<T, true>, 256, 0)
                                      : resident_ctas(mat_static_part_kernel<T, false>, 256, 0);
            const MatSegs sg = mat_segments(batch * bps, resident, no, 128);
            if (sg.nseg > 1) {
                const i64 blocks = batch * sg.nseg * bps;
                if (blocks > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: too many segments");
                void *ws = nullptr;
                keep_pool_memory();
                int rc = check_cuda(cudaMallocAsync(&ws, (size_t)batch * sg.nseg * 6 * np * sizeof(double), stream), "nbg_matrix: workspace");
                if (rc) return rc;
                if (corr)
                    mat_static_part_kernel<T, true><<<(unsigned)blocks, 256, 0, stream>>>(a, (double *)ws, batch, (int)nv, no, sg.seg_len, sg.nseg, bps);
                else
                    mat_static_part_kernel<T, false><<<(unsigned)blocks, 256, 0, stream>>>(a, (double *)ws, batch, (int)nv, no, sg.seg_len, sg.nseg, bps);
                rc = check_launch("nbg_matrix(static, partial sums)");
                if (!rc) {
                    const unsigned fb = (unsigned)((batch * np + 255) / 256);
                    if (corr)
                        mat_static_fold_kernel<T, true><<<fb, 256, 0, stream>>>((const double *)ws, out, batch, (int)nv, sg.nseg);
                    else
                        mat_static_fold_kernel<T, false><<<fb, 256, 0, stream>>>((const double *)ws, out, batch, (int)nv, sg.nseg);
                    rc = check_launch("nbg_matrix(static, fold)");
                }
                cudaFreeAsync(ws, stream);
                return rc;
            }
            if (corr)
                mat_static_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, out, batch, (int)nv, no);
            else
                mat_static_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, out, batch, (int)nv, no);
            break;
        }
        case NBG_MAT_MOVE_CORR:
        case NBG_MAT_MOVE_COV: {
            const bool corr = op == NBG_MAT_MOVE_CORR;
            const i64 np = nv * (nv + 1) / 2;
            const int bps = (int)((np + kMatSegThreads - 1) / kMatSegThreads);
            const bool table = window <= kMatRcpMax;
            // staged operands (NBG_MAT_STAGE=1, experiment): OB steps x nv variables per buffer, at most 1024
            // elements (4 per thread).  Measured SLOWER than the register prefetch of the unstaged form
            // (200 000 x 32 float64: 1.39 vs 0.81 ms): the chunk barrier makes every warp wait for the
            // slowest one and each step now starts with an exposed shared-memory round trip.
            int OB = (int)(1024 / nv);
            if (OB > 32) OB = 32;
            const bool stage = OB >= 8 && getenv("NBG_MAT_STAGE") != nullptr;
            const size_t tab_bytes = table ? (size_t)(((int)window + 2) & ~1) * sizeof(double) : 0;
            const size_t smem = tab_bytes + (stage ? (size_t)4 * OB * nv * sizeof(T) : 0);
            i64 resident = kNumSMs;
#define NBG_MAT_RES(C_, T_, S_)                                                                              \
    do {                                                                                                     \
        auto kern_ = mat_move_seg_kernel<T, C_, T_, S_>;                                                     \
        if (smem > ((size_t)48 << 10)) {                                                                     \
            const int rc_ = allow_big_smem(kern_, "nbg_matrix(move segments): cudaFuncSetAttribute");       \
            if (rc_) return rc_;                                                                             \
        }                                                                                                    \
        resident = resident_ctas(kern_, kMatSegThreads, smem);                                               \
    } while (0)
#define NBG_MAT_PICK(MAC)                                                   \
    do {                                                                    \
        if (corr) {                                                         \
            if (table) { if (stage) MAC(true, true, true); else MAC(true, true, false); }     \
            else { if (stage) MAC(true, false, true); else MAC(true, false, false); }         \
        } else {                                                            \
            if (table) { if (stage) MAC(false, true, true); else MAC(false, true, false); }   \
            else { if (stage) MAC(false, false, true); else MAC(false, false, false); }       \
        }                                                                   \
    } while (0)
            NBG_MAT_PICK(NBG_MAT_RES);
            const MatSegs sg = mat_segments(batch * bps, resident, no, window);
            if (sg.nseg > 1) {
                const i64 blocks = batch * sg.nseg * bps;
                if (blocks > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: too many segments");
#define NBG_MAT_SEG(C_, T_, S_)                                                                                              \
    mat_move_seg_kernel<T, C_, T_, S_><<<(unsigned)blocks, kMatSegThreads, smem, stream>>>(a, out, batch, no, (int)nv, window, \
                                                                                            min_count, sg.seg_len, sg.nseg, bps, OB)
                NBG_MAT_PICK(NBG_MAT_SEG);
#undef NBG_MAT_SEG
#undef NBG_MAT_RES
#undef NBG_MAT_PICK
            } else if (corr) {
                mat_move_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, out, batch, no, (int)nv, window, min_count);
            } else {
                mat_move_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, out, batch, no, (int)nv, window, min_count);
            }
            break;
        }
        case NBG_MAT_EXP_CORR:
        case NBG_MAT_EXP_COV: {
            const bool corr = op == NBG_MAT_EXP_CORR;
            const T *al = (const T *)alpha_;
            const i64 np = nv * (nv + 1) / 2;
            const int bps = (int)((np + kMatSegThreads - 1) / kMatSegThreads);
            const i64 resident = corr ? resident_ctas(mat_exp_seg_kernel<T, true, false>, kMatSegThreads, 0)
                                      : resident_ctas(mat_exp_seg_kernel<T, false, false>, kMatSegThreads, 0);
            const MatSegs sg = mat_segments(batch * bps, resident, no, 32);
            if (sg.nseg > 1) {
                const i64 blocks = batch * sg.nseg * bps;
                if (blocks > 0x7fffffff) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: too many segments");
                const size_t carry_bytes = (size_t)batch * sg.nseg * kMatExpStates * np * sizeof(double);
                const size_t decay_bytes = (size_t)batch * sg.nseg * 2 * sizeof(double);
                void *ws = nullptr;
                keep_pool_memory();
                int rc = check_cuda(cudaMallocAsync(&ws, carry_bytes + decay_bytes, stream), "nbg_matrix: workspace");
                if (rc) return rc;
                double *carry = (double *)ws, *decays = (double *)((char *)ws + carry_bytes);
#define NBG_MAT_EXP(C_, A_)                                                                                               \
    mat_exp_seg_kernel<T, C_, A_><<<(unsigned)blocks, kMatSegThreads, 0, stream>>>(a, al, alpha_per_item, (T)min_weight, out, \
                                                                                    batch, no, (int)nv, sg.seg_len, sg.nseg,  \
                                                                                    bps, carry, decays)
                if (corr) NBG_MAT_EXP(true, true); else NBG_MAT_EXP(false, true);
                rc = check_launch("nbg_matrix(exp, pass A)");
                const i64 ct = batch * kMatExpStates * np;
                if (!rc) mat_exp_carry_kernel<<<(unsigned)((ct + 255) / 256), 256, 0, stream>>>(carry, decays, batch, sg.nseg, np);
                if (!rc) rc = check_launch("nbg_matrix(exp, carry)");
                if (rc) {
                    cudaFreeAsync(ws, stream);
                    return rc;
                }
                if (corr) NBG_MAT_EXP(true, false); else NBG_MAT_EXP(false, false);
#undef NBG_MAT_EXP
                rc = check_launch("nbg_matrix");
                cudaFreeAsync(ws, stream);
                return rc;
            }
            if (corr)
                mat_exp_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, al, alpha_per_item, (T)min_weight, out, batch, no, (int)nv);
            else
                mat_exp_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, al, alpha_per_item, (T)min_weight, out, batch, no, (int)nv);
            break;
        }
        default: return fail(NBG_ERR_BAD_OP, "nbg_matrix: unknown op");
    }
    return check_launch("nbg_matrix");
}

}  // namespace
}  // namespace nbg

using namespace nbg;

extern "C" int nbg_matrix(int op, int dtype, const void *a, const void *alpha, int alpha_per_item, double min_weight,
                          void *out, int64_t batch, int64_t n_obs, int64_t n_vars, int64_t window, int64_t min_count,
                          void *stream) {
    if (op < NBG_MAT_NANCORR || op > NBG_MAT_EXP_COV) return fail(NBG_ERR_BAD_OP, "nbg_matrix: unknown op");
    if (dtype != NBG_F32 && dtype != NBG_F64) return fail(NBG_ERR_BAD_DTYPE, "nbg_matrix: dtype must be NBG_F32 or NBG_F64");
    if (batch < 0 || n_obs < 0 || n_vars < 0) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: negative extent");
    if (n_vars > 46340) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: n_vars^2 must fit in int32");
    if (batch == 0 || n_vars == 0) return NBG_OK;
    if (out == nullptr || (a == nullptr && n_obs > 0)) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: null pointer");
    const bool moving = op == NBG_MAT_MOVE_CORR || op == NBG_MAT_MOVE_COV;
    const bool expw = op == NBG_MAT_EXP_CORR || op == NBG_MAT_EXP_COV;
    if (moving && (window <= 0 || min_count < 0)) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: window must be positive, min_count >= 0");
    if (expw && alpha == nullptr && n_obs > 0) return fail(NBG_ERR_BAD_ARG, "nbg_matrix: alpha (one value per observation) is required");
    if (dtype == NBG_F32)
        return launch_matrix<float>(op, a, alpha, alpha_per_item, min_weight, out, batch, n_obs, n_vars, window, min_count,
                                    (cudaStream_t)stream);
    return launch_matrix<double>(op, a, alpha, alpha_per_item, min_weight, out, batch, n_obs, n_vars, window, min_count,
                                 (cudaStream_t)stream);
}
