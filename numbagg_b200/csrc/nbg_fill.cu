// nbg_fill.cu -- ffill / bfill for sm_100a.
//
// Replaces the per-slice loops of numbagg/funcs.py:294-326 (dispatched by ndfill,
// numbagg/decorators.py:417-487).  The loop carries (current value, lives_remaining); as a
// scan state that is (has_valid, bits of the last valid value, number of NaNs since it):
//   out[i] = a[i]                       if a[i] is valid
//          = last valid value           if `since <= limit`
//          = NaN (canonical quiet NaN)  otherwise / before the first valid value.
// Values are moved as raw bits, so the result is bit-exact with the reference.
//
// Row-tile path: generic chained scan (nbg_scan.cuh), one read + one write per element.
// bfill runs the same scan in mirrored order (REV).  Column-walk path (inner > 1): one
// thread per column, coalesced across columns, sequential along the core axis.
#include "nbg_scan.cuh"

namespace nbg {

struct FillAgg {
    int64_t has;            // 1 if the segment contains a valid value
    unsigned long long bits;  // raw bits of the last valid value
    int64_t dist;           // elements after the last valid value (segment length if none)
    __device__ static __forceinline__ FillAgg identity() { return FillAgg{0, 0ull, 0}; }
    __device__ static __forceinline__ FillAgg combine(const FillAgg &older, const FillAgg &newer) {
        if (newer.has) return newer;
        return FillAgg{older.has, older.bits, older.dist + newer.dist};
    }
    // a segment holding a valid value hides everything before it
    __device__ static __forceinline__ bool absorbing(const FillAgg &a) { return a.has != 0; }
};

template <typename T>
__device__ __forceinline__ unsigned long long to_bits(T v);
template <>
__device__ __forceinline__ unsigned long long to_bits<float>(float v) {
    return (unsigned long long)__float_as_uint(v);
}
template <>
__device__ __forceinline__ unsigned long long to_bits<double>(double v) {
    return (unsigned long long)__double_as_longlong(v);
}
template <typename T>
__device__ __forceinline__ T from_bits(unsigned long long b);
template <>
__device__ __forceinline__ float from_bits<float>(unsigned long long b) {
    return __uint_as_float((unsigned)b);
}
template <>
__device__ __forceinline__ double from_bits<double>(unsigned long long b) {
    return __longlong_as_double((long long)b);
}

template <typename T_, bool REV_>
struct FillPolicy {
    using T = T_;
    using Agg = FillAgg;
    static constexpr int NSTREAM = 1;
    static constexpr int MIN_CTAS = 6;  // <= 42 registers: six 35 KB tiles per SM
    static constexpr bool REV = REV_;
    static constexpr bool OVERLAP_INDEPENDENT = true;  // most chunks follow a valid value in the tile
    __device__ static __forceinline__ const T *stream_row(const ScanParams &p, int, int64_t row) {
        return reinterpret_cast<const T *>(p.in[0]) + row * p.n;
    }
    __device__ static __forceinline__ Agg load_carry(const ScanParams &p, int64_t row) {
        const int64_t *c = reinterpret_cast<const int64_t *>(p.carry_in) + row * NBG_FILL_STATE;
        return Agg{c[0], (unsigned long long)c[1], c[2]};
    }
    __device__ static __forceinline__ void store_agg(const ScanParams &p, int64_t row, const Agg &a) {
        int64_t *c = reinterpret_cast<int64_t *>(p.agg_out) + row * NBG_FILL_STATE;
        c[0] = a.has;
        c[1] = (int64_t)a.bits;
        c[2] = a.dist;
    }
    template <int E, class Get>
    __device__ static __forceinline__ Agg reduce(const ScanParams &, Get get, int cnt) {
        Agg a = Agg::identity();
#pragma unroll
        for (int k = 0; k < E; k++) {
            if (k >= cnt) break;
            const T v = get(0, k);
            if (is_nan(v)) {
                a.dist += 1;
            } else {
                a.has = 1;
                a.bits = to_bits(v);
                a.dist = 0;
            }
        }
        return a;
    }
    template <int E, class Get, class Put>
    __device__ static __forceinline__ void scan(const ScanParams &p, Agg st, Get get, Put put, int cnt) {
        const int64_t limit = p.limit;
#pragma unroll
        for (int k = 0; k < E; k++) {
            if (k >= cnt) break;
            const T v = get(0, k);
            if (is_nan(v)) {
                st.dist += 1;
                put(k, (st.has && st.dist <= limit) ? from_bits<T>(st.bits) : quiet_nan<T>());
            } else {
                st.has = 1;
                st.bits = to_bits(v);
                st.dist = 0;
                put(k, v);
            }
        }
    }
};

// ------------------------------------------------------------------------ column walk
struct FillColParams {
    const void *a;
    void *out;
    int64_t outer, n, inner, limit;
    const int64_t *carry_in;
    int64_t *agg_out;
    int rev;
};

template <typename T>
__global__ void __launch_bounds__(256) fill_colwalk_kernel(FillColParams p) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.outer * p.inner) return;
    const int64_t col = gid % p.inner, o = gid / p.inner;
    const T *a = reinterpret_cast<const T *>(p.a) + o * p.n * p.inner + col;
    T *out = p.out ? reinterpret_cast<T *>(p.out) + o * p.n * p.inner + col : nullptr;
    FillAgg st = FillAgg::identity();
    if (p.carry_in) {
        const int64_t *c = p.carry_in + gid * NBG_FILL_STATE;
        st = FillAgg{c[0], (unsigned long long)c[1], c[2]};
    }
    for (int64_t q = 0; q < p.n; q++) {
        const int64_t i = p.rev ? (p.n - 1 - q) : q;
        const T v = a[i * p.inner];
        T r;
        if (is_nan(v)) {
            st.dist += 1;
            r = (st.has && st.dist <= p.limit) ? from_bits<T>(st.bits) : quiet_nan<T>();
        } else {
            st.has = 1;
            st.bits = to_bits(v);
            st.dist = 0;
            r = v;
        }
        if (out) out[i * p.inner] = r;
    }
    if (p.agg_out) {
        int64_t *c = p.agg_out + gid * NBG_FILL_STATE;
        c[0] = st.has;
        c[1] = (int64_t)st.bits;
        c[2] = st.dist;
    }
}

// ---------------------------------------------------------------- sentinel patch (sharding)
// See include/nbg_b200.h: after a shard has been filled with the sentinel carry, only its leading
// (scan-order) run of sentinel values depends on the predecessors.  One CTA per 4096 scan
// positions; a CTA whose first position is not the sentinel lies beyond the run and leaves.
constexpr unsigned long long kFillSentinel64 = 0x7ff8dead5e171e1dull;
constexpr unsigned long long kFillSentinel32 = 0x7fc5e171ull;
template <typename T>
__device__ __forceinline__ unsigned long long fill_sentinel() {
    return sizeof(T) == 8 ? kFillSentinel64 : kFillSentinel32;
}
constexpr int kPatchTile = 4096;

template <typename T, bool REV>
__global__ void __launch_bounds__(256) fill_patch_kernel(T *__restrict__ out, int64_t n, int64_t tiles_per_row, int64_t limit,
                                                         const int64_t *__restrict__ carry) {
    const int64_t row = blockIdx.x / tiles_per_row;
    const int64_t c0 = (blockIdx.x % tiles_per_row) * (int64_t)kPatchTile;  // scan-order start
    T *o = out + row * n;
    auto at = [&](int64_t k) -> T * { return o + (REV ? (n - 1 - k) : k); };
    if (to_bits(*at(c0)) != fill_sentinel<T>()) return;  // the sentinel run is a prefix in scan order
    const int64_t has = carry[row * NBG_FILL_STATE + 0];
    const unsigned long long bits = (unsigned long long)carry[row * NBG_FILL_STATE + 1];
    const int64_t dist = carry[row * NBG_FILL_STATE + 2];
    for (int j = threadIdx.x; j < kPatchTile; j += 256) {
        const int64_t k = c0 + j;
        if (k >= n) break;
        T *p = at(k);
        if (to_bits(*p) != fill_sentinel<T>()) continue;
        *p = (has && dist + k + 1 <= limit) ? from_bits<T>(bits) : quiet_nan<T>();
    }
}

template <typename T>
struct FillTile;
template <>
struct FillTile<float> {
    static constexpr int THREADS = 256, E = 33;
};
template <>
struct FillTile<double> {
    static constexpr int THREADS = 256, E = 17;
};

template <typename T>
static int launch_fill(int dir, const void *a, void *out, int64_t outer, int64_t n, int64_t inner, int64_t limit,
                       const int64_t *carry_in, int64_t *agg_out, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (outer * n * inner == 0) return NBG_OK;
    if (inner == 1) {
        ScanParams p = {};
        p.in[0] = a;
        p.out = out;
        p.carry_in = carry_in;
        p.agg_out = agg_out;
        p.limit = limit;
        constexpr int TH = FillTile<T>::THREADS, E = FillTile<T>::E;
        if (dir == NBG_FFILL)
            return launch_scan_rowtile<FillPolicy<T, false>, TH, E>(p, outer, n, ws, ws_bytes, stream, "nbg_fill(ffill)");
        return launch_scan_rowtile<FillPolicy<T, true>, TH, E>(p, outer, n, ws, ws_bytes, stream, "nbg_fill(bfill)");
    }
    FillColParams p;
    p.a = a, p.out = out, p.outer = outer, p.n = n, p.inner = inner, p.limit = limit;
    p.carry_in = carry_in, p.agg_out = agg_out, p.rev = (dir == NBG_BFILL);
    const int64_t blocks = (outer * inner + 255) / 256;
    if (blocks > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_fill: grid too large");
    fill_colwalk_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return check_launch("nbg_fill(colwalk)");
}

}  // namespace nbg

extern "C" size_t nbg_fill_workspace_bytes(int itemsize, int64_t outer, int64_t n, int64_t inner) {
    using namespace nbg;
    if (inner != 1 || outer <= 0 || n <= 0) return 0;
    if (itemsize == 4)
        return scan_rowtile_workspace_bytes<FillPolicy<float, false>, FillTile<float>::THREADS, FillTile<float>::E>(outer, n);
    return scan_rowtile_workspace_bytes<FillPolicy<double, false>, FillTile<double>::THREADS, FillTile<double>::E>(outer, n);
}

extern "C" int nbg_fill(int dir, int itemsize, const void *a, void *out, int64_t outer, int64_t n, int64_t inner,
                        int64_t limit, const int64_t *carry_in, int64_t *agg_out, void *workspace,
                        size_t workspace_bytes, void *stream) {
    using namespace nbg;
    if (dir != NBG_FFILL && dir != NBG_BFILL) return fail(NBG_ERR_BAD_OP, "nbg_fill: dir must be NBG_FFILL or NBG_BFILL");
    if (outer < 0 || n < 0 || inner < 0) return fail(NBG_ERR_BAD_ARG, "nbg_fill: negative size");
    if (limit < 0) return fail(NBG_ERR_BAD_ARG, "nbg_fill: limit must be >= 0");
    if (outer * n * inner > 0 && !a) return fail(NBG_ERR_BAD_ARG, "nbg_fill: null input");
    if (!out && !agg_out) return fail(NBG_ERR_BAD_ARG, "nbg_fill: neither out nor agg_out given");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (itemsize == 4)
        return launch_fill<float>(dir, a, out, outer, n, inner, limit, carry_in, agg_out, workspace, workspace_bytes, st);
    if (itemsize == 8)
        return launch_fill<double>(dir, a, out, outer, n, inner, limit, carry_in, agg_out, workspace, workspace_bytes, st);
    return fail(NBG_ERR_BAD_DTYPE, "nbg_fill: itemsize must be 4 (float32) or 8 (float64)");
}

extern "C" uint64_t nbg_fill_sentinel_bits(int itemsize) {
    return itemsize == 8 ? nbg::kFillSentinel64 : nbg::kFillSentinel32;
}

extern "C" int nbg_fill_patch(int dir, int itemsize, void *out, int64_t outer, int64_t n, int64_t inner, int64_t limit,
                              const int64_t *carry, void *stream) {
    using namespace nbg;
    if (dir != NBG_FFILL && dir != NBG_BFILL) return fail(NBG_ERR_BAD_OP, "nbg_fill_patch: dir must be NBG_FFILL or NBG_BFILL");
    if (outer < 0 || n < 0 || inner < 0 || limit < 0) return fail(NBG_ERR_BAD_ARG, "nbg_fill_patch: negative argument");
    if (outer * n * inner == 0) return NBG_OK;
    if (inner != 1) return fail(NBG_ERR_UNSUPPORTED, "nbg_fill_patch: inner must be 1 (use the two-pass form otherwise)");
    if (!out || !carry) return fail(NBG_ERR_BAD_ARG, "nbg_fill_patch: null pointer");
    const int64_t tpr = (n + kPatchTile - 1) / kPatchTile;
    if (tpr * outer > INT32_MAX) return fail(NBG_ERR_UNSUPPORTED, "nbg_fill_patch: grid too large");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)(tpr * outer);
    if (itemsize == 4) {
        if (dir == NBG_FFILL) fill_patch_kernel<float, false><<<grid, 256, 0, st>>>(static_cast<float *>(out), n, tpr, limit, carry);
        else fill_patch_kernel<float, true><<<grid, 256, 0, st>>>(static_cast<float *>(out), n, tpr, limit, carry);
    } else if (itemsize == 8) {
        if (dir == NBG_FFILL) fill_patch_kernel<double, false><<<grid, 256, 0, st>>>(static_cast<double *>(out), n, tpr, limit, carry);
        else fill_patch_kernel<double, true><<<grid, 256, 0, st>>>(static_cast<double *>(out), n, tpr, limit, carry);
    } else {
        return fail(NBG_ERR_BAD_DTYPE, "nbg_fill_patch: itemsize must be 4 or 8");
    }
    return check_launch("nbg_fill_patch");
}
