// nbg_move_exp.cu -- exponentially weighted moving functions for sm_100a.
//
// Replaces the per-slice loops of numbagg/moving_exp.py:12-335 (dispatched by ndmoveexp,
// numbagg/decorators.py:344-414).  Every function is a first-order linear recurrence on a
// few channels,
//     s_c <- d_i * s_c (+ u_c(x_i) if x_i is valid),   d_i = 1 - alpha_i
// (sum_weight_2 decays by d_i^2, moving_exp.py:132), followed by a NON-linear read-out of
// the state (ratios, thresholds).  A segment is therefore summarised by the affine map
//     s -> D*s + U     with  D = prod d_i,  D2 = prod d_i^2,  U = state reached from zero,
// which composes associatively: (D_b, U_b) o (D_a, U_a) = (D_a*D_b, D_b*U_a + U_b).
// nbg_scan.cuh chains these maps across row tiles with decoupled look-back; outputs are
// always recomputed from the state (never patched), and the arithmetic inside a chunk is the
// reference's own sequence of roundings (decay first, then add; no FMA contraction).
//
// Algorithmic traffic: one read per input element + one write per output element; a scalar
// alpha costs nothing, a 1-D alpha is re-read per row from L2, an N-D alpha is a third stream.
#include <stdlib.h>

#include "nbg_scan.cuh"

namespace nbg {

// (fast_rcp / rcp_ok / qdiv / fdiv: nbg_common.cuh)

// -------------------------------------------------------------------------------------- ops
// GATE = false: the caller proved `weight >= min_weight` for every possible state (scalar alpha in
// [0, 1] and min_weight <= 0: the weight is a sum of non-negative terms), so the weight channel --
// always the LAST one -- is neither carried nor tested.
// ---- read-outs ---------------------------------------------------------------------------------
// Every read-out has ONE straight-line fast path -- reciprocals from the hardware seed, corrected
// quotients (nbg_common.cuh: they round exactly like the reference's divisions) -- followed by a single
// test that all divisors were inside the seed's range and every quotient came out finite; anything
// else (zero / subnormal / infinite sums) goes to an out-of-line copy of the reference's formula with
// IEEE divisions.  Inlining a division fallback behind each of the 4-5 quotients cost 130-250 bytes of
// spills per thread under the 64-register cap (config 3: move_exp_nanvar 7.9 -> 10.5 ms).
#define kExpInf __longlong_as_double(0x7ff0000000000000LL)
__device__ __forceinline__ double qd(double a, double b, double y) {  // unguarded corrected quotient
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}
__device__ __noinline__ double exp_var_ieee(double s0, double s1, double s2, double s3) {
    const double m = s1 / s2;
    const double var_biased = dsub(s0 / s2, dmul(m, m));
    const double bias = dsub(1.0, s3 / dmul(s2, s2));
    return bias > 0 ? var_biased / bias : __longlong_as_double(0x7ff8000000000000LL);
}
__device__ __noinline__ double exp_cov_ieee(double s0, double s1, double s2, double s3, double s4) {
    const double cov_biased = dsub(s2, dmul(s0, s1) / s3) / s3;
    const double bias = dsub(1.0, s4 / dmul(s3, s3));
    return bias > 0 ? cov_biased / bias : __longlong_as_double(0x7ff8000000000000LL);
}
__device__ __noinline__ double exp_corr_ieee(double s0, double s1, double s2, double s3, double s4, double s5, double s6) {
    const double cov = dsub(s2, dmul(s0, s1) / s3);
    const double var1 = dsub(s5, dmul(s0, s0) / s3);
    const double var2 = dsub(s6, dmul(s1, s1) / s3);
    const double bias = dsub(1.0, s4 / dmul(s3, s3));
    if (bias > 0) {
        const double den = sqrt(dmul(var1, var2));
        if (den > 0) return cov / den;
    }
    return __longlong_as_double(0x7ff8000000000000LL);
}

// contrib(): value added to each channel for one valid observation.  SQ_CH: the channel that
// decays with d^2 (-1: none).  output(): read-out from the state.
template <typename T, bool GATE = true>
struct ExpCount {  // moving_exp.py:12-38   channels: count, [weight]
    static constexpr int NIN = 1, NCH = GATE ? 2 : 1, SQ_CH = -1;
    static constexpr bool HAS_SEEN = false;
    __device__ static __forceinline__ void contrib(T, T, double alpha, double *c) {
        c[0] = 1.0;
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool, double mw) {
        return (!GATE || s[NCH - 1] >= mw) ? (T)s[0] : quiet_nan<T>();
    }
};
template <typename T, bool GATE = true>
struct ExpMean {  // moving_exp.py:41-72    channels: numer, denom, [weight]
    static constexpr int NIN = 1, NCH = GATE ? 3 : 2, SQ_CH = -1;
    static constexpr bool HAS_SEEN = false;
    __device__ static __forceinline__ void contrib(T a, T, double alpha, double *c) {
        c[0] = (double)a;
        c[1] = 1.0;
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool, double mw) {
        const bool pass = !GATE || s[NCH - 1] >= mw;
        if constexpr (std::is_same<T, float>::value) {
            // float32 output: divide in float32 when both operands are comfortably inside its
            // range (<= 1.5 ulp = 2e-7 from rounding the double quotient), else in double
            if (fabs(s[0]) < 1e30 && s[1] > 1e-30 && s[1] < 1e30)
                return pass ? __fdiv_rn((float)s[0], (float)s[1]) : quiet_nan<T>();
        }
        return pass ? (T)fdiv(s[0], s[1]) : quiet_nan<T>();
    }
};
template <typename T, bool GATE = true>
struct ExpSum {  // moving_exp.py:75-103   channels: numer, [weight]; NaN until the first valid
    static constexpr int NIN = 1, NCH = GATE ? 2 : 1, SQ_CH = -1;
    static constexpr bool HAS_SEEN = true;
    __device__ static __forceinline__ void contrib(T a, T, double alpha, double *c) {
        c[0] = (double)a;
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool seen, double mw) {
        return ((!GATE || s[NCH - 1] >= mw) && seen) ? (T)s[0] : quiet_nan<T>();
    }
};
template <typename T, bool SQRT, bool GATE = true>
struct ExpVar {  // moving_exp.py:106-224  channels: sum_x_2, sum_x, sum_weight, sum_weight_2, [weight]
    static constexpr int NIN = 1, NCH = GATE ? 5 : 4, SQ_CH = 3;
    static constexpr bool HAS_SEEN = false;
    __device__ static __forceinline__ void contrib(T a, T, double alpha, double *c) {
        c[0] = prod_as_input(a, a);
        c[1] = (double)a;
        c[2] = 1.0;
        c[3] = 1.0;
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool, double mw) {
        if (GATE && !(s[NCH - 1] >= mw)) return quiet_nan<T>();
        const double sw2 = dmul(s[2], s[2]);
        const double r = fast_rcp(s[2]);
        const double m = qd(s[1], s[2], r), e = qd(s[0], s[2], r);
        const double var_biased = dsub(e, dmul(m, m));
        const double t = qd(s[3], sw2, fast_rcp(sw2));
        const double bias = dsub(1.0, t);
        double v;
        if (rcp_ok(s[2]) && rcp_ok(sw2) && fabs(m) + fabs(e) + fabs(t) < kExpInf) {
            if (!(bias > 0)) return quiet_nan<T>();
            v = qd(var_biased, bias, fast_rcp(bias));
            if (!(rcp_ok(bias) && fabs(v) < kExpInf)) v = ieee_div(var_biased, bias);
        } else {
            v = exp_var_ieee(s[0], s[1], s[2], s[3]);
        }
        return (T)(SQRT ? sqrt(v) : v);  // (the seed + Newton root is not faster here: 7.71 vs 7.64 ms)
    }
};
template <typename T, bool GATE = true>
struct ExpCov {  // moving_exp.py:227-273  channels: sum_x1, sum_x2, sum_x1x2, sum_weight, sum_weight_2, [weight]
    static constexpr int NIN = 2, NCH = GATE ? 6 : 5, SQ_CH = 4;
    static constexpr bool HAS_SEEN = false;
    __device__ static __forceinline__ void contrib(T a, T b, double alpha, double *c) {
        c[0] = (double)a;
        c[1] = (double)b;
        c[2] = prod_as_input(a, b);
        c[3] = 1.0;
        c[4] = 1.0;
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool, double mw) {
        if (GATE && !(s[NCH - 1] >= mw)) return quiet_nan<T>();
        const double sw2 = dmul(s[3], s[3]);
        const double r = fast_rcp(s[3]);
        const double a = qd(dmul(s[0], s[1]), s[3], r);
        const double cov_biased = qd(dsub(s[2], a), s[3], r);
        const double t = qd(s[4], sw2, fast_rcp(sw2));
        const double bias = dsub(1.0, t);
        double v;
        if (rcp_ok(s[3]) && rcp_ok(sw2) && fabs(a) + fabs(cov_biased) + fabs(t) < kExpInf) {
            if (!(bias > 0)) return quiet_nan<T>();
            v = qd(cov_biased, bias, fast_rcp(bias));
            if (!(rcp_ok(bias) && fabs(v) < kExpInf)) v = ieee_div(cov_biased, bias);
        } else {
            v = exp_cov_ieee(s[0], s[1], s[2], s[3], s[4]);
        }
        return (T)v;
    }
};
template <typename T, bool GATE = true>
struct ExpCorr {  // moving_exp.py:276-335  + sum_x1_2, sum_x2_2  (the weight, when carried, is the last channel)
    static constexpr int NIN = 2, NCH = GATE ? 8 : 7, SQ_CH = 4;
    static constexpr bool HAS_SEEN = false;
    __device__ static __forceinline__ void contrib(T a, T b, double alpha, double *c) {
        c[0] = (double)a;
        c[1] = (double)b;
        c[2] = prod_as_input(a, b);
        c[3] = 1.0;
        c[4] = 1.0;
        c[5] = prod_as_input(a, a);
        c[6] = prod_as_input(b, b);
        if (GATE) c[NCH - 1] = alpha;
    }
    __device__ static __forceinline__ T output(const double *s, bool, double mw) {
        if (GATE && !(s[NCH - 1] >= mw)) return quiet_nan<T>();
        const double sw2 = dmul(s[3], s[3]);
        const double r = fast_rcp(s[3]);
        const double a = qd(dmul(s[0], s[1]), s[3], r), b = qd(dmul(s[0], s[0]), s[3], r), c = qd(dmul(s[1], s[1]), s[3], r);
        const double cov = dsub(s[2], a), var1 = dsub(s[5], b), var2 = dsub(s[6], c);
        const double t = qd(s[4], sw2, fast_rcp(sw2));
        const double bias = dsub(1.0, t);
        double v;
        if (rcp_ok(s[3]) && rcp_ok(sw2) && fabs(a) + fabs(b) + fabs(c) + fabs(t) < kExpInf) {
            if (!(bias > 0)) return quiet_nan<T>();
            const double vv = dmul(var1, var2);
            if (!(vv > 0)) return quiet_nan<T>();  // the reference's gate is sqrt(vv) > 0: the same set
            if (vv > 1e-35 && vv < 1e35) {
                v = dmul(cov, fast_rsqrt(vv));  // <= 2 ulp from cov / sqrt(vv)
            } else {
                const double den = ieee_sqrt(vv);
                if (!(den > 0)) return quiet_nan<T>();
                v = ieee_div(cov, den);
            }
        } else {
            v = exp_corr_ieee(s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
        }
        return (T)v;
    }
};

// -------------------------------------------------------------------------------- aggregate
// Words: D, [D2 if a channel decays by d^2], U[NCH], [seen if the op needs it].
template <class Op>
struct ExpAgg {
    static constexpr bool HAS_D2 = Op::SQ_CH >= 0;
    static constexpr int iD2 = 1;
    static constexpr int iU = 1 + (HAS_D2 ? 1 : 0);
    static constexpr int iSeen = iU + Op::NCH;
    static constexpr int NW = iSeen + (Op::HAS_SEEN ? 1 : 0);
    double w[NW];
    __device__ __forceinline__ double &D() { return w[0]; }
    __device__ __forceinline__ double D() const { return w[0]; }
    __device__ __forceinline__ double D2() const { return HAS_D2 ? w[iD2] : 1.0; }
    __device__ __forceinline__ void set_D2(double v) {
        if (HAS_D2) w[iD2] = v;
    }
    __device__ __forceinline__ double *U() { return w + iU; }
    __device__ __forceinline__ const double *U() const { return w + iU; }
    __device__ __forceinline__ bool seen() const { return Op::HAS_SEEN ? (w[Op::HAS_SEEN ? iSeen : 0] != 0.0) : false; }
    __device__ __forceinline__ void set_seen(bool v) {
        if (Op::HAS_SEEN) w[Op::HAS_SEEN ? iSeen : 0] = v ? 1.0 : 0.0;
    }
    __device__ static __forceinline__ ExpAgg identity() {
        ExpAgg a;
#pragma unroll
        for (int i = 0; i < NW; i++) a.w[i] = 0.0;
        a.w[0] = 1.0;
        a.set_D2(1.0);
        return a;
    }
    __device__ static __forceinline__ ExpAgg combine(const ExpAgg &older, const ExpAgg &newer) {
        ExpAgg r;
        r.w[0] = dmul(older.D(), newer.D());
        r.set_D2(dmul(older.D2(), newer.D2()));
#pragma unroll
        for (int c = 0; c < Op::NCH; c++)
            r.U()[c] = dadd(dmul(c == Op::SQ_CH ? newer.D2() : newer.D(), older.U()[c]), newer.U()[c]);
        r.set_seen(older.seen() || newer.seen());
        return r;
    }
    // Once the decay product of a segment has underflowed to exactly 0.0, nothing before the
    // segment can reach its end state (0 * finite == 0).  NOTE: a non-finite state (an inf
    // observation) more than ~log(2^-1074)/log(d) elements back would still dominate in the
    // reference (inf * d == inf); see DESIGN.md "known deviations".
    __device__ static __forceinline__ bool absorbing(const ExpAgg &a) {
        if (Op::HAS_SEEN && !a.seen()) return false;  // an all-NaN segment keeps the older flag
        return a.D() == 0.0 && (!HAS_D2 || a.D2() == 0.0);
    }
};

template <class Op, typename T>
__device__ __forceinline__ void exp_step(double *s, bool &seen, T a, T b, double alpha) {
    // decay first (moving_exp.py:60-63), then add the observation if valid
    const double d = dsub(1.0, alpha);
    const double d2 = dmul(d, d);
#pragma unroll
    for (int c = 0; c < Op::NCH; c++) s[c] = dmul(s[c], c == Op::SQ_CH ? d2 : d);
    const bool valid = Op::NIN == 2 ? !(is_nan(a) || is_nan(b)) : !is_nan(a);
    if (valid) {
        double u[Op::NCH];
        Op::contrib(a, b, alpha, u);
#pragma unroll
        for (int c = 0; c < Op::NCH; c++) s[c] = dadd(s[c], u[c]);
        seen = true;
    }
}

template <typename T_, class Op_, bool ALPHA_STREAM>
struct ExpPolicy {
    using T = T_;
    using Op = Op_;
    using Agg = ExpAgg<Op>;
    static constexpr int NSTREAM = Op::NIN + (ALPHA_STREAM ? 1 : 0);
    static constexpr int MIN_CTAS = (Op::NCH <= 3 && !ALPHA_STREAM) ? 5 : (Op::NCH <= 4 ? 4 : (Op::NCH <= 5 ? 3 : 2));  // register caps: 51 / 64 / 85 / 128 (measured, config 3)
    static constexpr bool REV = false;
    static constexpr bool OVERLAP_INDEPENDENT = false;  // a chunk is carry-free only after ~7k elements of decay
    __device__ static __forceinline__ const T *stream_row(const ScanParams &p, int s, int64_t row) {
        if (ALPHA_STREAM && s == Op::NIN)
            return reinterpret_cast<const T *>(p.in[2]) + (p.alpha_nd ? row * p.n : 0);
        return reinterpret_cast<const T *>(p.in[s]) + row * p.n;
    }
    __device__ static __forceinline__ Agg load_carry(const ScanParams &p, int64_t row) {
        const double *c = reinterpret_cast<const double *>(p.carry_in) + row * NBG_EXP_STATE;
        Agg a = Agg::identity();
#pragma unroll
        for (int q = 0; q < Op::NCH; q++) a.U()[q] = c[2 + q];
        a.set_seen(c[10] != 0.0);
        return a;
    }
    __device__ static __forceinline__ void store_agg(const ScanParams &p, int64_t row, const Agg &a) {
        double *c = reinterpret_cast<double *>(p.agg_out) + row * NBG_EXP_STATE;
        c[0] = a.D();
        c[1] = a.D2();
#pragma unroll
        for (int q = 0; q < 8; q++) c[2 + q] = q < Op::NCH ? a.U()[q < Op::NCH ? q : 0] : 0.0;
        c[10] = a.seen() ? 1.0 : 0.0;
    }
    template <int E, class Get>
    __device__ static __forceinline__ Agg reduce(const ScanParams &p, Get get, int cnt) {
        Agg a = Agg::identity();
        bool seen = false;
#pragma unroll
        for (int k = 0; k < E; k++) {
            if (k >= cnt) break;
            const T x = get(0, k);
            const T y = Op::NIN == 2 ? get(1, k) : x;
            const double alpha = ALPHA_STREAM ? (double)get(Op::NIN, k) : p.alpha_scalar;
            if (ALPHA_STREAM) {
                const double d = dsub(1.0, alpha);
                a.w[0] = dmul(a.D(), d);
                a.set_D2(dmul(a.D2(), dmul(d, d)));
            }
            exp_step<Op, T>(a.U(), seen, x, y, alpha);
        }
        if (!ALPHA_STREAM) {
            // scalar alpha: the chunk's decay product is d^cnt -- by squaring, not one multiply per element
            const double d = dsub(1.0, p.alpha_scalar);
            double pw = 1.0, b = d;
            for (int m = cnt; m > 0; m >>= 1) {
                if (m & 1) pw = dmul(pw, b);
                b = dmul(b, b);
            }
            a.w[0] = pw;
            a.set_D2(dmul(pw, pw));
        }
        a.set_seen(seen);
        return a;
    }
    template <int E, class Get, class Put>
    __device__ static __forceinline__ void scan(const ScanParams &p, Agg st, Get get, Put put, int cnt) {
        bool seen = st.seen();
        const double mw = p.min_weight;
#pragma unroll
        for (int k = 0; k < E; k++) {
            if (k >= cnt) break;
            const T x = get(0, k);
            const T y = Op::NIN == 2 ? get(1, k) : x;
            const double alpha = ALPHA_STREAM ? (double)get(Op::NIN, k) : p.alpha_scalar;
            exp_step<Op, T>(st.U(), seen, x, y, alpha);
            put(k, Op::output(st.U(), seen, mw));
        }
    }
};

// ------------------------------------------------------------------------ column walk
struct ExpColParams {
    const void *a1, *a2, *alpha;
    void *out;
    int64_t outer, n, inner;
    int alpha_mode;  // 0 scalar, 1 = 1-D over the core axis, 2 = full (outer, n, inner)
    double alpha_scalar, min_weight;
    const double *carry_in;
    double *agg_out;
};

template <typename T, class Op>
__global__ void __launch_bounds__(256) exp_colwalk_kernel(ExpColParams p) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.outer * p.inner) return;
    const int64_t col = gid % p.inner, o = gid / p.inner;
    const int64_t base = o * p.n * p.inner + col;
    const T *a = reinterpret_cast<const T *>(p.a1) + base;
    const T *b = Op::NIN == 2 ? reinterpret_cast<const T *>(p.a2) + base : a;
    const T *al = reinterpret_cast<const T *>(p.alpha);
    T *out = p.out ? reinterpret_cast<T *>(p.out) + base : nullptr;
    double s[Op::NCH];
    double D = 1.0, D2 = 1.0;
#pragma unroll
    for (int c = 0; c < Op::NCH; c++) s[c] = 0.0;
    bool seen = false;
    if (p.carry_in) {
        const double *c = p.carry_in + gid * NBG_EXP_STATE;
#pragma unroll
        for (int q = 0; q < Op::NCH; q++) s[q] = c[2 + q];
        seen = c[10] != 0.0;
    }
    for (int64_t i = 0; i < p.n; i++) {
        const T x = a[i * p.inner];
        const T y = Op::NIN == 2 ? b[i * p.inner] : x;
        double alpha = p.alpha_scalar;
        if (p.alpha_mode == 1) alpha = (double)al[i];
        if (p.alpha_mode == 2) alpha = (double)al[base + i * p.inner];
        const double d = dsub(1.0, alpha);
        D = dmul(D, d);
        D2 = dmul(D2, dmul(d, d));
        exp_step<Op, T>(s, seen, x, y, alpha);
        if (out) out[i * p.inner] = Op::output(s, seen, p.min_weight);
    }
    if (p.agg_out) {
        double *c = p.agg_out + gid * NBG_EXP_STATE;
        c[0] = D;
        c[1] = D2;
#pragma unroll
        for (int q = 0; q < 8; q++) c[2 + q] = 0.0;
#pragma unroll
        for (int q = 0; q < Op::NCH; q++) c[2 + q] = s[q];
        c[10] = seen ? 1.0 : 0.0;
    }
}

// ---------------------------------------------------------------------------------- launch
template <typename T>
struct ExpTile;
template <>
struct ExpTile<float> {
    static constexpr int THREADS = 256, E = 33;
};
template <>
struct ExpTile<double> {
    static constexpr int THREADS = 256, E = 17;
};

struct ExpArgs {
    const void *a1, *a2, *alpha;
    int alpha_nd;
    double alpha_scalar, min_weight;
    void *out;
    int64_t outer, n, inner;
    const double *carry_in;
    double *agg_out;
    void *ws;
    size_t ws_bytes;
    cudaStream_t stream;
};

template <typename T, class Op>
static int launch_exp(const ExpArgs &x) {
    if (x.outer * x.n * x.inner == 0) return NBG_OK;
    constexpr int TH = ExpTile<T>::THREADS, E = ExpTile<T>::E;
    if (x.inner == 1) {
        ScanParams p = {};
        p.in[0] = x.a1;
        p.in[1] = x.a2;
        p.in[2] = x.alpha;
        p.out = x.out;
        p.carry_in = x.carry_in;
        p.agg_out = x.agg_out;
        p.alpha_scalar = x.alpha_scalar;
        p.min_weight = x.min_weight;
        p.alpha_nd = x.alpha_nd;
        if (x.alpha)
            return launch_scan_rowtile<ExpPolicy<T, Op, true>, TH, E>(p, x.outer, x.n, x.ws, x.ws_bytes, x.stream,
                                                                      "nbg_move_exp(rowtile, alpha array)");
        return launch_scan_rowtile<ExpPolicy<T, Op, false>, TH, E>(p, x.outer, x.n, x.ws, x.ws_bytes, x.stream,
                                                                   "nbg_move_exp(rowtile)");
    }
    ExpColParams p;
    p.a1 = x.a1, p.a2 = x.a2, p.alpha = x.alpha, p.out = x.out;
    p.outer = x.outer, p.n = x.n, p.inner = x.inner;
    p.alpha_mode = x.alpha ? (x.alpha_nd ? 2 : 1) : 0;
    p.alpha_scalar = x.alpha_scalar, p.min_weight = x.min_weight;
    p.carry_in = x.carry_in, p.agg_out = x.agg_out;
    const int64_t blocks = (x.outer * x.inner + 255) / 256;
    if (blocks > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_move_exp: grid too large");
    exp_colwalk_kernel<T, Op><<<(unsigned)blocks, 256, 0, x.stream>>>(p);
    return check_launch("nbg_move_exp(colwalk)");
}

template <typename T, bool GATE>
static int dispatch_exp_gate(int op, const ExpArgs &x) {
    switch (op) {
        case NBG_EXP_NANCOUNT:
            return launch_exp<T, ExpCount<T, GATE>>(x);
        case NBG_EXP_NANMEAN:
            return launch_exp<T, ExpMean<T, GATE>>(x);
        case NBG_EXP_NANSUM:
            return launch_exp<T, ExpSum<T, GATE>>(x);
        case NBG_EXP_NANVAR:
            return launch_exp<T, ExpVar<T, false, GATE>>(x);
        case NBG_EXP_NANSTD:
            return launch_exp<T, ExpVar<T, true, GATE>>(x);
        case NBG_EXP_NANCOV:
            return launch_exp<T, ExpCov<T, GATE>>(x);
        case NBG_EXP_NANCORR:
            return launch_exp<T, ExpCorr<T, GATE>>(x);
        default:
            return fail(NBG_ERR_BAD_OP, "nbg_move_exp: unknown op");
    }
}

template <typename T>
static int dispatch_exp(int op, const ExpArgs &x) {
    // weight = sum of alpha * d^k over valid observations >= 0 whenever 0 <= alpha <= 1, so the gate
    // `weight >= min_weight` (moving_exp.py:35, 69, 98, 151, 269, 322) is always open for
    // min_weight <= 0 -- the default -- and the weight channel need not be carried at all
    const bool no_gate = x.alpha == nullptr && x.alpha_scalar >= 0.0 && x.alpha_scalar <= 1.0 && x.min_weight <= 0.0 &&
                         !getenv("NBG_EXP_GATE");
    return no_gate ? dispatch_exp_gate<T, false>(op, x) : dispatch_exp_gate<T, true>(op, x);
}

template <typename T>
static size_t exp_ws_bytes(int op, int64_t outer, int64_t n) {
    constexpr int TH = ExpTile<T>::THREADS, E = ExpTile<T>::E;
    // the largest aggregate (corr) bounds every op; size by op to keep small ops small
    switch (op) {
        case NBG_EXP_NANCOUNT:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpCount<T>, false>, TH, E>(outer, n);
        case NBG_EXP_NANMEAN:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpMean<T>, false>, TH, E>(outer, n);
        case NBG_EXP_NANSUM:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpSum<T>, false>, TH, E>(outer, n);
        case NBG_EXP_NANVAR:
        case NBG_EXP_NANSTD:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpVar<T, false>, false>, TH, E>(outer, n);
        case NBG_EXP_NANCOV:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpCov<T>, false>, TH, E>(outer, n);
        default:
            return scan_rowtile_workspace_bytes<ExpPolicy<T, ExpCorr<T>, false>, TH, E>(outer, n);
    }
}

}  // namespace nbg

extern "C" size_t nbg_move_exp_workspace_bytes(int op, int dtype, int64_t outer, int64_t n, int64_t inner) {
    using namespace nbg;
    if (inner != 1 || outer <= 0 || n <= 0) return 0;
    return dtype == NBG_F32 ? exp_ws_bytes<float>(op, outer, n) : exp_ws_bytes<double>(op, outer, n);
}

extern "C" int nbg_move_exp(int op, int dtype, const void *a1, const void *a2, const void *alpha, int alpha_nd,
                            double alpha_scalar, double min_weight, void *out, int64_t outer, int64_t n,
                            int64_t inner, const double *carry_in, double *agg_out, void *workspace,
                            size_t workspace_bytes, void *stream) {
    using namespace nbg;
    if (outer < 0 || n < 0 || inner < 0) return fail(NBG_ERR_BAD_ARG, "nbg_move_exp: negative size");
    const bool two = (op == NBG_EXP_NANCOV || op == NBG_EXP_NANCORR);
    if (outer * n * inner > 0 && (!a1 || (two && !a2))) return fail(NBG_ERR_BAD_ARG, "nbg_move_exp: null input");
    if (!out && !agg_out) return fail(NBG_ERR_BAD_ARG, "nbg_move_exp: neither out nor agg_out given");
    ExpArgs x;
    x.a1 = a1, x.a2 = a2, x.alpha = alpha, x.alpha_nd = alpha_nd, x.alpha_scalar = alpha_scalar;
    x.min_weight = min_weight, x.out = out, x.outer = outer, x.n = n, x.inner = inner;
    x.carry_in = carry_in, x.agg_out = agg_out, x.ws = workspace, x.ws_bytes = workspace_bytes;
    x.stream = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case NBG_F32:
            return dispatch_exp<float>(op, x);
        case NBG_F64:
            return dispatch_exp<double>(op, x);
        default:
            return fail(NBG_ERR_BAD_DTYPE, "nbg_move_exp: dtype must be NBG_F32 or NBG_F64");
    }
}
