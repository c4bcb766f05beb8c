"""Test helper: walks the packed 8-byte slot tables produced by the C++ packer (fmc_pack.hpp) in
NumPy, so the specialiser can be checked on CPU against the oracle.  Test code, not product code."""
import numpy as np

FLAGS = (3, 12, 13, 14, 16)


def sim_rows(num17, zero_missing, scaler=None):
    """[n,15] float32 feature rows of the simulation preset (row 14 = -inf)."""
    n = num17.shape[0]
    x = np.array(num17, dtype=np.float64, copy=True)
    if scaler is not None:
        cols, mean, scale = scaler
        for j, k in enumerate(cols):
            x[:, k] = (x[:, k] - mean[j]) / scale[j]
    v = x.astype(np.float32)
    rows = np.zeros((n, 15), dtype=np.float32)
    src = [0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 16]
    for r, k in enumerate(src):
        if k < v.shape[1]:
            rows[:, r] = v[:, k]
    if zero_missing:
        for r, rb in ((1, 11), (2, 12), (4, 13)):
            z = rows[:, r] == 0
            rows[:, rb] = np.where(z, np.inf, rows[:, r])
            rows[:, r] = np.where(z, -np.inf, rows[:, r])
    rows[:, 14] = -np.inf
    return rows


def predict_rows(num17, zero_missing, scaler=None, xgb=True):
    """[n,30] float32 feature rows of the predict preset: 17 numerics (A views), the B views of the 12 non-flags,
    row 29 = -inf.  xgboost forests: a NaN (and, for the CSR-fed boosters, an exact zero) is missing -> A view -inf,
    B view +inf; a missing flag is an absent one (0)."""
    n = num17.shape[0]
    x = np.array(num17, dtype=np.float64, copy=True)
    if scaler is not None:
        cols, mean, scale = scaler
        for j, k in enumerate(cols):
            x[:, k] = (x[:, k] - mean[j]) / scale[j]
    v = np.zeros((n, 17), dtype=np.float32)
    v[:, :x.shape[1]] = x.astype(np.float32)
    rows = np.zeros((n, 30), dtype=np.float32)
    rows[:, :17] = v
    if xgb:
        nb = 17
        for k in range(17):
            if k in FLAGS:
                rows[:, k] = np.where(np.isnan(v[:, k]), 0.0, v[:, k])
                continue
            z = np.isnan(v[:, k]) | ((v[:, k] == 0) if zero_missing else False)
            rows[:, k] = np.where(z, -np.inf, v[:, k])
            rows[:, nb] = np.where(z, np.inf, v[:, k])
            nb += 1
    rows[:, 29] = -np.inf
    return rows


CHILD_MASK = 0x000FFFF8
FEAT_BYTES = 128


def walk(slots, stream, consts, meta, rows, skl, base):
    """Margins [n, n_outputs] from the packed tables (fmc_pack.hpp), walked exactly like the kernels do:
    groups of `ilp` trees, D branch-free levels each, constants added at their place in the tree order."""
    n = rows.shape[0]
    lo = (slots & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (slots >> np.uint64(32)).astype(np.uint32)
    ilp = meta["ilp"]
    out = np.zeros((n, meta["n_outputs"]), dtype=np.float64)
    ar = np.arange(n)
    # player mode appends 0/1 rows (one per usage entry) behind the -inf row
    assert rows.shape[1] >= meta["ninf_row"] + 1 and np.all(np.isneginf(rows[:, meta["ninf_row"]]))

    def value(w_lo, w_hi):
        if skl:
            return ((w_hi.astype(np.uint64) << np.uint64(32)) | w_lo.astype(np.uint64)).view(np.float64)
        return w_lo.view(np.float32)

    for o in range(meta["n_outputs"]):
        acc = np.full(n, base[o], dtype=np.float64 if skl else np.float32)
        sp = meta["stream_off"][o]
        cp = meta["consts_off"][o]
        for g in range(meta["n_groups"][o]):
            roots = stream[sp + g * ilp: sp + (g + 1) * ilp]
            r_lo = (roots & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            r_hi = (roots >> np.uint64(32)).astype(np.uint32)
            depth = int(r_hi[0] & 7) | (int(r_hi[1] & 1) << 3)
            has_consts = bool(r_hi[1] & 2)
            window = int(r_hi[2] & 7) | (int(r_hi[3] & 7) << 3)      # 1 MiB window of the group inside the table
            cur_lo = np.tile(r_lo, (n, 1))
            cur_hi = np.tile(r_hi, (n, 1))
            for _ in range(depth):
                for q in range(ilp):
                    frow = (cur_hi[:, q] >> np.uint32(20)) // np.uint32(FEAT_BYTES)
                    assert np.all((cur_hi[:, q] >> np.uint32(20)) % np.uint32(FEAT_BYTES) == 0)
                    fv = rows[ar, frow]
                    thr = cur_lo[:, q].view(np.float32)
                    right = ~(fv <= thr) if skl else ~(fv < thr)
                    a = (window << 20) + (cur_hi[:, q] & np.uint32(CHILD_MASK)).astype(np.int64) + 8 * right
                    assert np.all(a % 8 == 0) and a.max() < 8 * len(slots)
                    cur_lo[:, q] = lo[a // 8]
                    cur_hi[:, q] = hi[a // 8]
            counts = [0] * ilp
            if has_consts:
                cw = int(consts[cp]); cp += 1
                counts = [(cw >> (8 * q)) & 0xFF for q in range(ilp)]
            for q in range(ilp):
                for _ in range(counts[q]):
                    w = np.array([consts[cp]], dtype=np.uint64); cp += 1
                    c = value((w & np.uint64(0xFFFFFFFF)).astype(np.uint32), (w >> np.uint64(32)).astype(np.uint32))[0]
                    acc = acc + c
                leaf = value(np.ascontiguousarray(cur_lo[:, q]), np.ascontiguousarray(cur_hi[:, q]))
                if not skl:     # a lane must be sitting on a self-pointing leaf
                    assert np.all(cur_hi[:, q] >> np.uint32(20) == meta["ninf_row"] * FEAT_BYTES)
                acc = acc + leaf
        out[:, o] = acc
    return out
