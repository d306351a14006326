/*
 * nbg_b200.h -- C ABI of the B200-native numbagg hot path (libnbg_b200.so).
 *
 * Drop-in boundary.  numbagg has no native code: its native boundary is the NumPy gufunc
 * inner loop that numba emits for each decorated body,
 *     void loop(char **args, npy_intp *dims, npy_intp *steps, void *data)
 * (numba/np/ufunc/wrappers.py:334-343), built at numbagg/decorators.py:151-162 (move / exp),
 * :427-463 (fill) and :533-556 (grouped).  Each entry point below replaces one family of
 * those loops: `dims[0]` (outer-loop count) and the core dimension become explicit
 * (outer, n, inner) sizes, `steps` are implied by C-contiguity, scalars are passed by value
 * and operands are DEVICE pointers.  INTEGRATION.md shows the ctypes binding a numbagg
 * maintainer would add in decorators.py.
 *
 * Conventions
 *  - All data pointers are device pointers (cudaMalloc / torch CUDA tensors), never host.
 *  - Every array operand is the C-contiguous 3-D view (outer, n, inner) of the caller's
 *    array with the core axis in the middle; axis=-1 of a C-contiguous array is inner == 1,
 *    any other axis of a C-contiguous array is inner > 1 -- no copy is ever needed for a
 *    C- or F-contiguous input.  Output has the same view.
 *  - Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *    stream).  No allocation, no host synchronisation; the caller owns every buffer and
 *    the kernels write every output element (outputs may be uninitialised on entry).
 *  - Return value: NBG_OK or a negative nbg_status; hot-path kernels never raise on data
 *    (NaN, inf, /0 are values).  nbg_last_error() gives a thread-local message.
 *  - No CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef NBG_B200_H
#define NBG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBG_ABI_VERSION 1

typedef enum {
    NBG_OK = 0,
    NBG_ERR_BAD_DTYPE = -1,
    NBG_ERR_BAD_OP = -2,
    NBG_ERR_BAD_ARG = -3,
    NBG_ERR_UNSUPPORTED = -4, /* legal in numbagg but not built yet (see message) */
    NBG_ERR_CUDA = -5,        /* launch / runtime failure: see nbg_last_error() */
    NBG_ERR_WORKSPACE = -6    /* workspace too small */
} nbg_status;

typedef enum { NBG_F32 = 0, NBG_F64 = 1, NBG_I32 = 2, NBG_I64 = 3 } nbg_dtype;

/* numbagg/moving.py:12-275 */
typedef enum {
    NBG_MOVE_MEAN = 0,
    NBG_MOVE_SUM = 1,
    NBG_MOVE_STD = 2,
    NBG_MOVE_VAR = 3,
    NBG_MOVE_COV = 4, /* two inputs */
    NBG_MOVE_CORR = 5 /* two inputs */
} nbg_move_op;

/* numbagg/moving_exp.py:12-335 */
typedef enum {
    NBG_EXP_NANCOUNT = 0,
    NBG_EXP_NANMEAN = 1,
    NBG_EXP_NANSUM = 2,
    NBG_EXP_NANVAR = 3,
    NBG_EXP_NANSTD = 4,
    NBG_EXP_NANCOV = 5, /* two inputs */
    NBG_EXP_NANCORR = 6 /* two inputs */
} nbg_exp_op;

/* numbagg/funcs.py:294-326 */
typedef enum { NBG_FFILL = 0, NBG_BFILL = 1 } nbg_fill_dir;

/* numbagg/grouped.py:7-270 */
typedef enum {
    NBG_GROUP_NANMEAN = 0,
    NBG_GROUP_NANSUM = 1,
    NBG_GROUP_NANCOUNT = 2,
    NBG_GROUP_NANARGMAX = 3,
    NBG_GROUP_NANARGMIN = 4,
    NBG_GROUP_NANFIRST = 5,
    NBG_GROUP_NANLAST = 6,
    NBG_GROUP_NANPROD = 7,
    NBG_GROUP_NANSUM_OF_SQUARES = 8,
    NBG_GROUP_NANVAR = 9,
    NBG_GROUP_NANSTD = 10,
    NBG_GROUP_NANMIN = 11,
    NBG_GROUP_NANMAX = 12,
    NBG_GROUP_NANANY = 13,
    NBG_GROUP_NANALL = 14
} nbg_group_op;

int nbg_abi_version(void);
const char *nbg_last_error(void);
/* Number of kernels this library has launched in the calling process (all families). */
int64_t nbg_launch_count(void);

/*
 * Moving-window functions.  Replaces the gufunc loops "(a),(),()->(a)" /
 * "(a),(a),(),()->(a)" built by ndmove (numbagg/decorators.py:275-341) over the bodies in
 * numbagg/moving.py.  dtype: NBG_F32 | NBG_F64 (accumulation is always double, like the
 * reference).  b: second input for COV/CORR, else NULL.  Requires 0 < window; min_count >= 0
 * (the per-op max(min_count, 1|2) clamp of moving.py:18,127,158,193,232 is applied inside).
 *
 * Core-axis sharding (multi-GPU): when this call covers only positions [g0, g0+n) of a
 * longer core axis, pass in a_halo/b_halo the `halo_len` elements that precede the shard
 * (view (outer, halo_len, inner), normally halo_len = min(window, g0)); NULL / 0 otherwise.
 */
int nbg_move(int op, int dtype, const void *a, const void *b, void *out, int64_t outer,
             int64_t n, int64_t inner, int64_t window, int64_t min_count, const void *a_halo,
             const void *b_halo, int64_t halo_len, void *stream);

/*
 * Exponential moving functions.  Replaces the loops "(a),(a),()->(a)" /
 * "(a),(a),(a),()->(a)" built by ndmoveexp (numbagg/decorators.py:344-414) over
 * numbagg/moving_exp.py.  alpha: device array of the SAME dtype as the data, either 1-D of
 * length n shared by every slice (alpha_nd == 0) or the full (outer, n, inner) view
 * (alpha_nd == 1); alpha == NULL means the scalar `alpha_scalar` for every position
 * (decorators.py:398-400 broadcasts it; here it costs no memory traffic).
 *
 * State exchange for core-axis sharding: NBG_EXP_STATE doubles per slice,
 *   [0]=prod(decay) [1]=prod(decay^2) [2..9]=channel values [10]=seen-any-valid flag.
 * carry_in  (nullable): state at the start of this shard (only [2..10] are read).
 * agg_out   (nullable): this shard's aggregate computed from a ZERO carry; composing
 *           aggregates left to right with s' = D*s + U gives the carry of the next shard.
 * out may be NULL when only agg_out is wanted (aggregate-only pass, no output traffic).
 */
#define NBG_EXP_STATE 11
int nbg_move_exp(int op, int dtype, const void *a1, const void *a2, const void *alpha,
                 int alpha_nd, double alpha_scalar, double min_weight, void *out,
                 int64_t outer, int64_t n, int64_t inner, const double *carry_in,
                 double *agg_out, void *workspace, size_t workspace_bytes, void *stream);
size_t nbg_move_exp_workspace_bytes(int op, int dtype, int64_t outer, int64_t n, int64_t inner);

/*
 * ffill / bfill.  Replaces the loop "(n),()->(n)" built by ndfill
 * (numbagg/decorators.py:417-487) over numbagg/funcs.py:294-326.  itemsize: 4 or 8 bytes of
 * an IEEE float dtype (integers have no NaN: the caller copies them).  limit >= 0.
 * State exchange for core-axis sharding: NBG_FILL_STATE int64 words per slice,
 *   [0]=has_valid [1]=raw bits of the last valid value [2]=distance from that value to the
 *   end of the shard (= shard length when has_valid == 0).  Same carry_in / agg_out / out
 *   rules as nbg_move_exp; for bfill "start" and "end" are mirrored.
 */
#define NBG_FILL_STATE 3
int nbg_fill(int dir, int itemsize, const void *a, void *out, int64_t outer, int64_t n,
             int64_t inner, int64_t limit, const int64_t *carry_in, int64_t *agg_out,
             void *workspace, size_t workspace_bytes, void *stream);
size_t nbg_fill_workspace_bytes(int itemsize, int64_t outer, int64_t n, int64_t inner);
/*
 * Single-pass core-axis sharding of ffill / bfill (one read + one write per element instead of
 * an aggregate pass followed by a scan pass).  Step 1: nbg_fill() with carry_in = the SENTINEL
 * carry {1, nbg_fill_sentinel_bits(itemsize), 0}: the result is final everywhere except in the
 * leading (bfill: trailing) NaN run of the shard, which now holds the sentinel -- a NaN payload
 * that no output can otherwise contain, because inputs that are NaN are never copied.  Step 2,
 * after the shards' aggregates have been exchanged and folded into `carry` ((outer, 3) int64,
 * same words as carry_in): nbg_fill_patch() rewrites the sentinel run in place -- the carried
 * value while `dist + position + 1 <= limit`, NaN afterwards.  It reads one element per 4096
 * outside that run.  inner must be 1.
 */
uint64_t nbg_fill_sentinel_bits(int itemsize);
int nbg_fill_patch(int dir, int itemsize, void *out, int64_t outer, int64_t n, int64_t inner,
                   int64_t limit, const int64_t *carry, void *stream);

/*
 * Grouped reductions.  Replaces the loops "(a..),(a..),(z)" / "(a..),(a..),(),(z)" built by
 * groupndreduce (numbagg/decorators.py:490-674) over numbagg/grouped.py.  values is the
 * (rows, n) C-contiguous matrix obtained after the dispatcher's moveaxis (decorators.py:639,
 * 653) with the core dims flattened in C order; labels is (n,) shared by every row
 * (labels_per_row == 0) or (rows, n); out is (rows, num_labels) in the values dtype.
 * label < 0 or label >= num_labels => element skipped (the reference corrupts memory on
 * the latter; we guard).  vdtype: any nbg_dtype (NANMEAN/NANVAR/NANSTD floats only -- the
 * dispatcher casts integers to float64 first, decorators.py:613-619); ldtype: NBG_I32|I64.
 *
 * Three-step form for element-sharded inputs (multi-GPU): init -> accumulate (any number
 * of shards, `index_offset` = flat index of the shard's first element) -> [caller combines
 * workspaces across devices] -> finalize.  nbg_group() runs the three steps on one device.
 * The workspace holds one record of nbg_group_record_words(op) 8-byte slots per (row, label),
 * ws[row][label][slot] (slot meanings per op: DESIGN.md "group workspace"), followed by
 * per-call scratch (the column plan
 * of the shared-label kernel); nbg_group_workspace_bytes() sizes both for shards of up to
 * `n` elements per row.  Only the first record_words*rows*num_labels*8 bytes (after rounding
 * the base up to 256) are state that must be exchanged between devices.
 */
#define NBG_GROUP_WS_CHANNELS 4 /* upper bound of nbg_group_record_words() */
int nbg_group_record_words(int op); /* 8-byte slots per (row, label) record: 1, 2 or 4 */
/* Where the channels of an op's record live for a (rows, num_labels) table, so that a caller can
 * run collectives on single channels: layout[0] = stride between records in 8-byte words,
 * layout[1..3] = offset of channel 0..2 in 8-byte words from the (256-byte aligned) state base,
 * layout[4] = 1 when channels are separate planes (offsets are multiples of rows*num_labels).
 * Channels per op: DESIGN.md "group workspace". */
int nbg_group_record_layout(int op, int64_t rows, int64_t num_labels, int64_t layout[5]);
size_t nbg_group_workspace_bytes(int op, int vdtype, int64_t rows, int64_t n, int64_t num_labels);
int nbg_group_init(int op, int vdtype, void *workspace, int64_t rows, int64_t num_labels,
                   void *stream);
int nbg_group_accumulate(int op, int vdtype, int ldtype, const void *values,
                         const void *labels, int labels_per_row, void *workspace,
                         size_t workspace_bytes, int64_t rows, int64_t n, int64_t num_labels,
                         int64_t index_offset, void *stream);
/* Merge workspace `other` (accumulated over LATER elements) into `accum`; order matters for
 * first/last and for arg* ties.  Used to fold all-gathered per-device partials. */
int nbg_group_combine(int op, int vdtype, void *accum, const void *other, int64_t rows,
                      int64_t num_labels, void *stream);
int nbg_group_finalize(int op, int vdtype, const void *workspace, void *out, int64_t rows,
                       int64_t num_labels, int64_t ddof, void *stream);
int nbg_group(int op, int vdtype, int ldtype, const void *values, const void *labels,
              int labels_per_row, void *out, int64_t rows, int64_t n, int64_t num_labels,
              int64_t ddof, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Plain NaN-aware reductions (the first row past the hot path): allnan, anynan,
 * nancount, nansum, nanmean, nanvar, nanstd behind ndaggregate (numbagg/decorators.py:188-260)
 * and nanargmax, nanargmin, nanmax, nanmin behind ndreduce (decorators.py:906-1031), bodies
 * in numbagg/funcs.py:23-242.  `a` is a C-contiguous (outer, n, inner) array and the middle
 * axis is reduced (the dispatchers' move_axes/moveaxis + flatten of the reduced axes, which
 * the caller does by choosing outer/n/inner or by permuting); `out` has outer*inner
 * elements of:  uint8 (ALLNAN, ANYNAN) | int64 (NANCOUNT, NANARGMAX, NANARGMIN) | the input
 * dtype (NANSUM; NANMEAN, NANVAR, NANSTD -- float inputs only; NANMAX, NANMIN on floats) |
 * int64 (NANMAX, NANMIN on ints, as numba types them).  NANARG* write -1 for a slice without
 * a non-NaN element and NANMAX/NANMIN write NaN; the reference raises ValueError for the
 * former and for n == 0 -- that check belongs to the caller.  ddof: NANVAR / NANSTD only.
 *
 * Element-sharded form (multi-GPU): nbg_reduce_partial writes one record of
 * NBG_REDUCE_STATE_WORDS 8-byte words per output, states[word][j] (index_offset = position
 * of the shard's first element along the reduced axis, used by NANARG*); the caller gathers
 * the records of all shards into states[part][word][j] and nbg_reduce_merge folds and
 * finalizes them (n_total: reduced length over all shards, for ANYNAN).
 */
typedef enum {
    NBG_RED_ALLNAN = 0,
    NBG_RED_ANYNAN = 1,
    NBG_RED_NANCOUNT = 2,
    NBG_RED_NANSUM = 3,
    NBG_RED_NANMEAN = 4,
    NBG_RED_NANVAR = 5,
    NBG_RED_NANSTD = 6,
    NBG_RED_NANARGMAX = 7,
    NBG_RED_NANARGMIN = 8,
    NBG_RED_NANMAX = 9,
    NBG_RED_NANMIN = 10
} nbg_reduce_op;
#define NBG_REDUCE_STATE_WORDS 3
size_t nbg_reduce_workspace_bytes(int op, int dtype, int64_t outer, int64_t n, int64_t inner);
int nbg_reduce(int op, int dtype, const void *a, void *out, int64_t outer, int64_t n,
               int64_t inner, int64_t ddof, void *workspace, size_t workspace_bytes,
               void *stream);
int nbg_reduce_partial(int op, int dtype, const void *a, void *states, int64_t outer,
                       int64_t n, int64_t inner, int64_t index_offset, void *workspace,
                       size_t workspace_bytes, void *stream);
int nbg_reduce_merge(int op, int dtype, const void *states, int64_t parts, int64_t outs,
                     void *out, int64_t n_total, int64_t ddof, void *stream);

/*
 * nanquantile / nanmedian (numbagg/funcs.py:245-291, 332-335 behind
 * ndquantile, numbagg/decorators.py:821-901).  `a` is a C-contiguous (rows, n) float64
 * matrix (the dispatcher's move_axes puts the reduced axes last; other dtypes are cast to
 * float64 like NumPy does for the reference's only loop), `q` holds m <= 16 quantiles in
 * [0, 1] (NaN allowed -> NaN result) ON THE DEVICE, `out` is (rows, m) float64.  NaN = missing
 * value; a row without data gives NaN.  Results are bit-identical to the reference (exact
 * selection + its interpolation arithmetic).  Rows longer than 4096 elements need a workspace.
 */
size_t nbg_quantile_workspace_bytes(int64_t rows, int64_t n, int64_t m);
int nbg_quantile(const void *a, const void *q, void *out, int64_t rows, int64_t n, int64_t m,
                 void *workspace, size_t workspace_bytes, void *stream);

/*
 * Pairwise-complete covariance / correlation matrices: nancorrmatrix,
 * nancovmatrix (numbagg/funcs.py:338-532 behind ndmatrix, decorators.py:677-740),
 * move_corrmatrix, move_covmatrix (numbagg/moving_matrix.py:16-204 behind ndmovematrix,
 * decorators.py:743-818) and move_exp_nancorrmatrix, move_exp_nancovmatrix
 * (moving_matrix.py:207-432 behind ndmoveexpmatrix, decorators.py:1031-1100).
 * Static ops: `a` is C-contiguous (batch, n_vars, n_obs), `out` (batch, n_vars, n_vars).
 * Moving / exponential ops: `a` is (batch, n_obs, n_vars), `out` (batch, n_obs, n_vars,
 * n_vars); `window`, `min_count` for the moving ops; for the exponential ops `alpha` holds one
 * decay weight per observation in the dtype of `a` -- (n_obs) shared by all batch items
 * (alpha_per_item = 0) or (batch, n_obs) -- and `min_weight` gates the output.
 * dtype: NBG_F32 | NBG_F64.  With n_obs < 8192 results are bit-identical to numbagg (same
 * operations, same order, running sums in the input dtype).  Longer observation axes are cut
 * into segments that run in parallel (windows rebuilt from the preceding observations,
 * exponential states carried as affine maps, static sums folded in order): results then agree
 * with numbagg to the rounding of its running sums, NaN masks exactly.  The exponential and the
 * static forms take their carry / partial-sum workspace from the device's stream-ordered memory
 * pool (cudaMallocAsync on `stream`).
 */
typedef enum {
    NBG_MAT_NANCORR = 0,
    NBG_MAT_NANCOV = 1,
    NBG_MAT_MOVE_CORR = 2,
    NBG_MAT_MOVE_COV = 3,
    NBG_MAT_EXP_CORR = 4,
    NBG_MAT_EXP_COV = 5
} nbg_matrix_op;
int nbg_matrix(int op, int dtype, const void *a, const void *alpha, int alpha_per_item,
               double min_weight, void *out, int64_t batch, int64_t n_obs, int64_t n_vars,
               int64_t window, int64_t min_count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NBG_B200_H */
