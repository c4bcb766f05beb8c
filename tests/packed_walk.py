"""Test helper: walks the packed 8-byte slot tables produced by the C++ packer (fmc_pack.hpp) in
NumPy, so the specialiser can be checked on CPU against the oracle.  Test code, not product code."""
import numpy as np

FLAGS = (3, 12, 13, 14, 16)


def sim_rows(num17, zero_missing, scaler=None):
    """[n,14] float32 feature rows of the simulation preset."""
    n = num17.shape[0]
    x = np.array(num17, dtype=np.float64, copy=True)
    if scaler is not None:
        cols, mean, scale = scaler
        for j, k in enumerate(cols):
            x[:, k] = (x[:, k] - mean[j]) / scale[j]
    v = x.astype(np.float32)
    rows = np.zeros((n, 14), dtype=np.float32)
    src = [0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 16]
    for r, k in enumerate(src):
        if k < v.shape[1]:
            rows[:, r] = v[:, k]
    if zero_missing:
        for r, rb in ((1, 11), (2, 12), (4, 13)):
            z = rows[:, r] == 0
            rows[:, rb] = np.where(z, np.inf, rows[:, r])
            rows[:, r] = np.where(z, -np.inf, rows[:, r])
    return rows


def predict_rows(num17, zero_missing, scaler=None):
    n = num17.shape[0]
    x = np.array(num17, dtype=np.float64, copy=True)
    if scaler is not None:
        cols, mean, scale = scaler
        for j, k in enumerate(cols):
            x[:, k] = (x[:, k] - mean[j]) / scale[j]
    v = np.zeros((n, 17), dtype=np.float32)
    v[:, :x.shape[1]] = x.astype(np.float32)
    rows = np.zeros((n, 29), dtype=np.float32)
    rows[:, :17] = v
    if zero_missing:
        nb = 17
        for k in range(17):
            if k in FLAGS:
                continue
            z = v[:, k] == 0
            rows[:, k] = np.where(z, -np.inf, v[:, k])
            rows[:, nb] = np.where(z, np.inf, v[:, k])
            nb += 1
    return rows


def walk(slots, roots, meta, rows, skl, feat_bits, base):
    """Margins [n, n_outputs] from packed tables; same accumulation order as the kernels."""
    n = rows.shape[0]
    lo = (slots & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (slots >> np.uint64(32)).astype(np.uint32)
    internal = hi.view(np.int32) >= 0x50000000
    row_of = ((hi >> np.uint32(20)) & np.uint32(0xFF)) // np.uint32(4)
    child = hi & np.uint32((1 << 20) - 1)
    thr = lo.view(np.float32)
    leaf64 = slots.view(np.float64)
    leaf32 = lo.view(np.float32)
    out = np.zeros((n, meta["n_outputs"]), dtype=np.float64)
    rp = meta["rounds_padded"]
    ar = np.arange(n)
    for o in range(meta["n_outputs"]):
        acc = np.full(n, base[o], dtype=np.float64 if skl else np.float32)
        for t in range(rp):
            # the walk starts from the inline COPY of the root slot (roots[o][t] = lo, hi)
            r_lo, r_hi = roots[(o * rp + t) * 2], roots[(o * rp + t) * 2 + 1]
            root = np.uint64(r_lo) | (np.uint64(r_hi) << np.uint64(32))
            hits = np.flatnonzero(slots == root)
            assert hits.size, "inline root slot is not a copy of a table slot"
            idx = np.full(n, hits[0], dtype=np.int64)
            live = internal[idx]
            while live.any():
                i = idx[live]
                fv = rows[ar[live], row_of[i]]
                right = ~(fv <= thr[i]) if skl else ~(fv < thr[i])
                idx[live] = child[i].astype(np.int64) + right
                live = internal[idx]
            acc = acc + (leaf64[idx] if skl else leaf32[idx])
        out[:, o] = acc
    return out
