"""CPU model of the scan filter's survivor bookkeeping (csrc/scan_epilogue.cuh): variant 0 (per-lane mask loop) vs
variant 3 (warp-cooperative extraction).  Both must produce the same stash contents, counters and overflow-pass
writes for every warp; run before flipping OM_SCAN_VARIANT.   python tools/sim_scan_variants.py"""
import numpy as np

K, C = 8, 4096  # kStash, list capacity


def variant0(vals, t, lim_cols):
    L, NC, _ = vals.shape
    k, n, stash = np.zeros(L, int), np.zeros(L, int), [[] for _ in range(L)]
    for c in range(NC):
        for l in range(L):
            for i in range(32):
                if c * 32 + i < lim_cols and vals[l, c, i] > t[l]:
                    if k[l] < K:
                        stash[l].append((c * 32 + i, vals[l, c, i]))
                        k[l] += 1
                    n[l] += 1
    out = [dict() for _ in range(L)]
    if (n > K).any():  # second pass: lanes with more than K survivors append the excess at their reserved position
        skip, pos2 = np.full(L, K), np.where(n > K, 100, C)
        for c in range(NC):
            for l in range(L):
                if n[l] <= K:
                    continue
                for i in range(32):
                    if c * 32 + i < lim_cols and vals[l, c, i] > t[l]:
                        if skip[l] > 0:
                            skip[l] -= 1
                        else:
                            if pos2[l] < C:
                                out[l][pos2[l]] = (c * 32 + i, vals[l, c, i])
                            pos2[l] += 1
    return stash, k, n, out


def variant3(vals, t, lim_cols):
    L, NC, _ = vals.shape
    k, n, stash = np.zeros(L, int), np.zeros(L, int), [dict() for _ in range(L)]

    def sweep(pass_, skip=None, pos2=None, out=None):
        for c in range(NC):
            col0 = c * 32
            if col0 >= lim_cols:
                continue
            lim = lim_cols - col0
            mx = vals[:, c, :].max(axis=1)
            hot = [l for l in range(L) if not (pass_ == 1 and n[l] <= K) and mx[l] > t[l]]
            for lh in hot:  # lane i holds column i of lane lh's chunk
                x = vals[lh, c, :]
                sv = (x > t[lh]) & (np.arange(32) < lim)
                cnt, rank = int(sv.sum()), np.cumsum(sv) - sv
                if pass_ == 0:
                    kl = k[lh]
                    for i in range(32):
                        if sv[i] and kl + rank[i] < K:
                            stash[lh][kl + rank[i]] = (col0 + i, x[i])
                    k[lh], n[lh] = min(K, kl + cnt), n[lh] + cnt
                else:
                    sk, p0 = skip[lh], pos2[lh]
                    for i in range(32):
                        if sv[i] and rank[i] >= sk and p0 + (rank[i] - sk) < C:
                            out[lh][p0 + (rank[i] - sk)] = (col0 + i, x[i])
                    used = min(cnt, sk)
                    skip[lh], pos2[lh] = sk - used, p0 + (cnt - used)

    sweep(0)
    out = [dict() for _ in range(L)]
    if (n > K).any():
        sweep(1, np.full(L, K), np.where(n > K, 100, C), out)
    return [[stash[l][j] for j in sorted(stash[l])] for l in range(L)], k, n, out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for trial in range(300):
        vals = rng.standard_normal((32, 4, 32)).astype(np.float32)
        t = rng.choice([0.5, 1.5, 2.5, -1.0, 9.0], size=32).astype(np.float32)
        lim_cols = int(rng.choice([128, 100, 97, 33]))
        a, b = variant0(vals, t, lim_cols), variant3(vals, t, lim_cols)
        assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all() and a[3] == b[3], trial
    print("variant 3 bookkeeping == variant 0 on 300 random warps (overflow pass and ragged last tile included)")
