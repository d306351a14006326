"""CPU stand-in for the CUDA kernels, used ONLY by tests/test_distributed_cpu.py to exercise
numbagg_b200.distributed's exchange logic (halo assembly, carry folding, partial-state
combination) over a 2-rank gloo group on machines without a GPU.  Window functions go through
the oracle; the carry/aggregate protocols of include/nbg_b200.h (NBG_EXP_STATE,
NBG_FILL_STATE, group workspace channels) are restated here in numpy for small inputs."""

import numpy as np
import torch

from oracle import oracle

_EXP = {
    # name: (n_inputs, contributions(x, y, alpha) -> list, squared-decay channel, has_seen)
    "move_exp_nancount": (1, lambda x, y, a: [1.0, a], None),
    "move_exp_nanmean": (1, lambda x, y, a: [x, 1.0, a], None),
    "move_exp_nansum": (1, lambda x, y, a: [x, a], None),
    "move_exp_nanvar": (1, lambda x, y, a: [x * x, x, 1.0, 1.0, a], 3),
    "move_exp_nanstd": (1, lambda x, y, a: [x * x, x, 1.0, 1.0, a], 3),
    "move_exp_nancov": (2, lambda x, y, a: [x, y, x * y, 1.0, 1.0, a], 4),
    "move_exp_nancorr": (2, lambda x, y, a: [x, y, x * y, 1.0, 1.0, a, x * x, y * y], 4),
}


def _exp_output(name, s, seen, mw):
    nan = np.nan
    with np.errstate(all="ignore"):
        if name == "move_exp_nancount":
            return s[0] if s[1] >= mw else nan
        if name == "move_exp_nanmean":
            return s[0] / s[1] if s[2] >= mw else nan
        if name == "move_exp_nansum":
            return s[0] if (s[1] >= mw and seen) else nan
        if name in ("move_exp_nanvar", "move_exp_nanstd"):
            m = s[1] / s[2]
            vb = s[0] / s[2] - m * m
            bias = 1 - s[3] / (s[2] * s[2])
            if s[4] >= mw and bias > 0:
                v = vb / bias
                return np.sqrt(v) if name.endswith("std") else v
            return nan
        if name == "move_exp_nancov":
            cb = (s[2] - s[0] * s[1] / s[3]) / s[3]
            bias = 1 - s[4] / (s[3] * s[3])
            return cb / bias if (s[5] >= mw and bias > 0) else nan
        cov = s[2] - s[0] * s[1] / s[3]
        v1 = s[6] - s[0] * s[0] / s[3]
        v2 = s[7] - s[1] * s[1] / s[3]
        bias = 1 - s[4] / (s[3] * s[3])
        if s[5] >= mw and bias > 0:
            den = np.sqrt(v1 * v2)
            return cov / den if den > 0 else nan
        return nan


def _slices_last(t, axis):
    a = np.moveaxis(t.numpy(), axis, -1)
    return np.ascontiguousarray(a).reshape(-1, a.shape[-1]), a.shape


# record layout per op (ws_layout() in numbagg_b200/csrc/nbg_group.cu): (words, slot of ch0/ch1/ch2)
def _layout(name):
    if name == "group_nancount":
        return 1, (0, 0, 0)
    if name == "group_nanmean":
        return 2, (0, 0, 1)
    if name in ("group_nanvar", "group_nanstd"):
        return 4, (0, 1, 2)
    if name in ("group_nanfirst", "group_nanlast", "group_nanargmax", "group_nanargmin"):
        return 2, (0, 1, 0)
    return 1, (0, 0, 0)


def _channels(rec, name):
    words, slots = _layout(name)
    return [rec[..., slots[0]], rec[..., slots[1]], rec[..., slots[2]]]


class OracleBackend:
    @staticmethod
    def move(name, arrs, window, min_count, axis, halos):
        full = [torch.cat([h, a], dim=axis) if halos else a for a, h in zip(arrs, halos or [None] * len(arrs))]
        out = getattr(oracle, name)(*[f.numpy() for f in full], window=window, min_count=min_count, axis=axis)
        h = halos[0].shape[axis] if halos else 0
        return torch.from_numpy(np.ascontiguousarray(out)).narrow(axis, h, arrs[0].shape[axis]).contiguous()

    @staticmethod
    def move_exp(name, arrs, alpha, min_weight, axis, carry_in, want_agg, want_out):
        nin, contrib, sq = _EXP[name]
        xs, shape = _slices_last(arrs[0], axis)
        ys = _slices_last(arrs[1], axis)[0] if nin == 2 else xs
        S, n = xs.shape
        if torch.is_tensor(alpha):
            al = alpha.numpy()
            al = np.broadcast_to(al, (S, n)) if al.ndim <= 1 else _slices_last(alpha, axis)[0]
        else:
            al = np.full((S, n), float(alpha))
        out = np.full((S, n), np.nan)
        agg = np.zeros((S, 11))
        for r in range(S):
            st = np.zeros(8)
            seen = False
            D = D2 = 1.0
            if carry_in is not None:
                st = carry_in[r, 2:10].numpy().copy()
                seen = bool(carry_in[r, 10] != 0)
            for i in range(n):
                a = float(al[r, i])
                d = 1.0 - a
                dec = np.full(8, d)
                if sq is not None:
                    dec[sq] = d * d
                st = st * dec
                D *= d
                D2 *= d * d
                x, y = float(xs[r, i]), float(ys[r, i])
                if not (np.isnan(x) or np.isnan(y)):
                    c = contrib(x, y, a)
                    st[: len(c)] += c
                    seen = True
                out[r, i] = _exp_output(name, st, seen, min_weight)
            agg[r] = np.concatenate([[D, D2], st, [1.0 if seen else 0.0]])
        o = torch.from_numpy(np.ascontiguousarray(np.moveaxis(out.reshape(shape), -1, axis))) if want_out else None
        return o, (torch.from_numpy(agg) if want_agg else None)

    @staticmethod
    def fill(name, t, limit, axis, carry_in, want_agg, want_out):
        xs, shape = _slices_last(t, axis)
        S, n = xs.shape
        rev = name == "bfill"
        out = np.full((S, n), np.nan)
        agg = np.zeros((S, 3), dtype=np.int64)
        for r in range(S):
            has, bits, dist = (0, 0, 0) if carry_in is None else [int(v) for v in carry_in[r]]
            for q in range(n):
                i = n - 1 - q if rev else q
                v = xs[r, i]
                if np.isnan(v):
                    dist += 1
                    out[r, i] = np.array([bits], dtype=np.int64).view(np.float64)[0] if (has and dist <= limit) else np.nan
                else:
                    has, bits, dist = 1, int(np.array([v]).view(np.int64)[0]), 0
                    out[r, i] = v
            agg[r] = (has, bits, dist)
        o = torch.from_numpy(np.ascontiguousarray(np.moveaxis(out.reshape(shape), -1, axis))) if want_out else None
        return o, (torch.from_numpy(agg) if want_agg else None)

    SENTINEL = 0x7FF8DEAD5E171E1D  # any NaN payload no output can otherwise hold

    @staticmethod
    def fill_sentinel(itemsize):
        return OracleBackend.SENTINEL

    @staticmethod
    def fill_patch(name, out, limit, axis, carry):
        """In-place rewrite of the leading (scan-order) sentinel run from the folded carry."""
        a = out.numpy()
        nd = a.ndim
        assert axis % nd == nd - 1
        flat = a.reshape(-1, a.shape[-1])
        bits = flat.view(np.int64)
        rev = name == "bfill"
        for r in range(flat.shape[0]):
            has, cb, dist = [int(v) for v in carry[r]]
            n = flat.shape[1]
            for q in range(n):
                i = n - 1 - q if rev else q
                if bits[r, i] != OracleBackend.SENTINEL:
                    break
                bits[r, i] = cb if (has and dist + q + 1 <= limit) else np.array([np.nan]).view(np.int64)[0]

    @staticmethod
    def group_channels(name, state, rows, K):
        words, slots = _layout(name)
        flat = state.reshape(-1)
        return [torch.as_strided(flat, (rows * K,), (words,), slots[c]) for c in range(3)]

    # ---- group workspace protocol (numbagg_b200/csrc/nbg_group.cu header), float64 values
    @staticmethod
    def _key(v):
        v = np.where(v == 0.0, 0.0, v)
        b = v.view(np.uint64)
        return np.where(b >> np.uint64(63) != 0, ~b, b | (np.uint64(1) << np.uint64(63)))

    @staticmethod
    def group_partial(name, values, labels, K, index_offset, labels_per_row):
        v = values.numpy().astype(np.float64)
        rows, n = v.shape
        lab = labels.numpy().reshape(-1, n) if labels_per_row else np.broadcast_to(labels.numpy(), (rows, n))
        rec = np.zeros((rows, K, _layout(name)[0]), dtype=np.int64)
        st = _channels(rec, name)
        f0 = st[0].view(np.float64)
        f1 = st[1].view(np.float64)
        k0 = st[0].view(np.uint64)
        if name == "group_nanprod":
            f0[:] = 1.0
        if name == "group_nanall":
            st[0][:] = 1
        if name in ("group_nanfirst", "group_nanargmax", "group_nanargmin"):
            st[1][:] = np.iinfo(np.int64).max
        if name == "group_nanlast":
            st[1][:] = -1
        for r in range(rows):
            for i in range(n):
                l, x = int(lab[r, i]), v[r, i]
                if l < 0 or l >= K or np.isnan(x):
                    continue
                gi = index_offset + i
                if name in ("group_nansum", "group_nanmean", "group_nanvar", "group_nanstd"):
                    f0[r, l] += x
                if name in ("group_nanvar", "group_nanstd"):
                    f1[r, l] += x * x
                if name == "group_nansum_of_squares":
                    f0[r, l] += x * x
                if name in ("group_nanmean", "group_nancount", "group_nanvar", "group_nanstd"):
                    st[2][r, l] += 1
                if name == "group_nanprod":
                    f0[r, l] *= x
                if name == "group_nanany" and x != 0:
                    st[0][r, l] = 1
                if name == "group_nanall" and x == 0:
                    st[0][r, l] = 0
                if name == "group_nanfirst" and gi < st[1][r, l]:
                    st[1][r, l] = gi
                    f0[r, l] = x
                if name == "group_nanlast" and gi > st[1][r, l]:
                    st[1][r, l] = gi
                    f0[r, l] = x
                if name in ("group_nanmax", "group_nanmin", "group_nanargmax", "group_nanargmin"):
                    k = OracleBackend._key(np.array([x]))[0]
                    if name in ("group_nanmin", "group_nanargmin"):
                        k = ~k
                    if k > k0[r, l]:
                        k0[r, l] = k
                        st[1][r, l] = gi if "arg" in name else st[1][r, l]
        return torch.from_numpy(rec)

    @staticmethod
    def group_combine(name, vdtype, acc, other):
        ra, ro = acc.numpy(), other.numpy()
        a = _channels(ra, name)
        o = _channels(ro, name)
        ak, ok = a[0].view(np.uint64), o[0].view(np.uint64)
        if name == "group_nanprod":
            a[0].view(np.float64)[:] *= o[0].view(np.float64)
        elif name in ("group_nanmin", "group_nanmax"):
            np.maximum(ak, ok, out=ak)
        elif name == "group_nanany":
            a[0][:] |= o[0]
        elif name == "group_nanall":
            a[0][:] &= o[0]
        elif name == "group_nanfirst":
            m = o[1] < a[1]
            a[1][m], a[0][m] = o[1][m], o[0][m]
        elif name == "group_nanlast":
            m = o[1] > a[1]
            a[1][m], a[0][m] = o[1][m], o[0][m]
        elif name in ("group_nanargmax", "group_nanargmin"):
            gt = ok > ak
            eq = (ok == ak) & (o[1] < a[1])
            a[1][gt | eq] = o[1][gt | eq]
            ak[gt] = ok[gt]
        else:
            raise AssertionError(name)

    @staticmethod
    def group_finalize(name, vdtype, state, ddof):
        rec = state.numpy()
        st = _channels(rec, name)
        f0, f1, cnt = st[0].view(np.float64), st[1].view(np.float64), st[2]
        k0 = st[0].view(np.uint64)
        with np.errstate(all="ignore"):
            if name in ("group_nansum", "group_nansum_of_squares", "group_nanprod"):
                out = f0.copy()
            elif name == "group_nancount":
                out = cnt.astype(np.float64)
            elif name in ("group_nanany", "group_nanall"):
                out = st[0].astype(np.float64)
            elif name == "group_nanmean":
                out = np.where(cnt == 0, np.nan, f0 / cnt)
            elif name in ("group_nanvar", "group_nanstd"):
                den = cnt - ddof
                q = (f1 - f0 * f0 / cnt) / den
                out = np.where(den <= 0, np.nan, np.sqrt(q) if name.endswith("std") else q)
            elif name in ("group_nanmax", "group_nanmin"):
                k = k0 if name == "group_nanmax" else ~k0
                b = np.where(k >> np.uint64(63) != 0, k & ~(np.uint64(1) << np.uint64(63)), ~k)
                out = np.where(k0 == 0, np.nan, b.view(np.float64))
            elif name in ("group_nanfirst", "group_nanlast"):
                none = (st[1] == np.iinfo(np.int64).max) if name == "group_nanfirst" else (st[1] < 0)
                out = np.where(none, np.nan, f0)
            else:
                out = np.where(k0 == 0, np.nan, st[1].astype(np.float64))
        return torch.from_numpy(out.astype(vdtype))

    # ---- plain-reduction state records (include/nbg_b200.h: NBG_REDUCE_STATE_WORDS), float64
    @staticmethod
    def reduce_partial(name, shard, axes, index_offset):
        (axis,) = axes
        a = np.moveaxis(shard.numpy().astype(np.float64), axis, -1)
        bshape = a.shape[:-1]
        flat = a.reshape(-1, a.shape[-1])
        outs, n = flat.shape
        st = np.zeros((3, outs), dtype=np.int64)
        f1, f2 = st[1].view(np.float64), st[2].view(np.float64)
        for j in range(outs):
            x = flat[j]
            ok = ~np.isnan(x)
            v = x[ok]
            if name in ("allnan", "anynan", "nancount"):
                st[0, j] = ok.sum()
            elif name == "nansum":
                st[0, j] = np.array([v.sum()]).view(np.int64)[0]
            elif name == "nanmean":
                st[0, j] = ok.sum()
                f1[j] = v.sum()
            elif name in ("nanvar", "nanstd"):
                st[0, j] = ok.sum()
                f1[j] = v.mean() if v.size else 0.0
                f2[j] = ((v - v.mean()) ** 2).sum() if v.size else 0.0
            elif name in ("nanmax", "nanmin"):
                st[0, j] = 1 if v.size else 0
                f1[j] = (v.max() if name == "nanmax" else v.min()) if v.size else (-np.inf if name == "nanmax" else np.inf)
            elif name in ("nanargmax", "nanargmin"):
                if v.size:
                    i = int(np.nanargmax(x) if name == "nanargmax" else np.nanargmin(x))
                    st[0, j] = index_offset + i
                    f1[j] = x[i]
                else:
                    st[0, j] = -1
                    f1[j] = -np.inf if name == "nanargmax" else np.inf
            else:
                raise AssertionError(name)
        return torch.from_numpy(st), (lambda out: out.reshape(bshape))

    @staticmethod
    def reduce_merge(name, vdtype, states, n_total, ddof):
        st = states.numpy()
        parts, _, outs = st.shape
        w0 = st[:, 0, :]
        f1, f2 = st[:, 1, :].view(np.float64), st[:, 2, :].view(np.float64)
        with np.errstate(all="ignore"):
            if name == "allnan":
                out = w0.sum(0) == 0
            elif name == "anynan":
                out = w0.sum(0) < n_total
            elif name == "nancount":
                out = w0.sum(0)
            elif name == "nansum":
                out = st[:, 0, :].view(np.float64).sum(0)
            elif name == "nanmean":
                c = w0.sum(0)
                out = np.where(c > 0, f1.sum(0) / c, np.nan)
            elif name in ("nanvar", "nanstd"):
                c = np.zeros(outs)
                mean = np.zeros(outs)
                m2 = np.zeros(outs)
                for p in range(parts):  # Chan's update, as in nbg_reduce.cu
                    cb = w0[p].astype(np.float64)
                    tot = np.where(c + cb > 0, c + cb, 1.0)
                    d = f1[p] - mean
                    m2 = np.where(cb > 0, m2 + f2[p] + d * d * c * cb / tot, m2)
                    mean = np.where(cb > 0, mean + d * cb / tot, mean)
                    c = c + cb
                out = np.where(c > ddof, m2 / (c - ddof), np.nan)
                if name == "nanstd":
                    out = np.sqrt(out)
            elif name in ("nanmax", "nanmin"):
                ext = (np.max if name == "nanmax" else np.min)(f1, axis=0)
                out = np.where(w0.sum(0) > 0, ext, np.nan)
            else:  # nanargmax / nanargmin: best key, smallest index among ties
                out = np.full(outs, -1, dtype=np.int64)
                key = np.full(outs, -np.inf if name == "nanargmax" else np.inf)
                for p in range(parts):
                    better = (f1[p] > key) if name == "nanargmax" else (f1[p] < key)
                    take = (w0[p] >= 0) & ((out < 0) | better | ((f1[p] == key) & (w0[p] < out)))
                    out = np.where(take, w0[p], out)
                    key = np.where(take, f1[p], key)
        return torch.from_numpy(np.asarray(out))
