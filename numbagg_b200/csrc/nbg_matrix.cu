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
        case NBG_MAT_NANCORR: mat_static_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, out, batch, (int)nv, no); break;
        case NBG_MAT_NANCOV: mat_static_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, out, batch, (int)nv, no); break;
        case NBG_MAT_MOVE_CORR:
            mat_move_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, out, batch, no, (int)nv, window, min_count);
            break;
        case NBG_MAT_MOVE_COV:
            mat_move_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, out, batch, no, (int)nv, window, min_count);
            break;
        case NBG_MAT_EXP_CORR:
            mat_exp_kernel<T, true><<<grid, kMatThreads, 0, stream>>>(a, (const T *)alpha_, alpha_per_item, (T)min_weight, out,
                                                                    batch, no, (int)nv);
            break;
        case NBG_MAT_EXP_COV:
            mat_exp_kernel<T, false><<<grid, kMatThreads, 0, stream>>>(a, (const T *)alpha_, alpha_per_item, (T)min_weight, out,
                                                                     batch, no, (int)nv);
            break;
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
