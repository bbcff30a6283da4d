"""CPU restatement of HParm's qualifier expansion for whole utterances -- TEST INFRASTRUCTURE ONLY
(only tests/ may import this; the product path is htk_b200/csrc/hfb_feat.cuh).

Follows HTKLib/HParm.c:1618-1722 (AddQualifiers), :1552-1599 (AddDiffs, tables: hdValid = tlValid = 0,
not V1COMPAT), HTKLib/HSigP.c:827-857 (Regress) and HSigP.c:803-823 (FZeroMean), in float32 with the reference's
order of operations.  Pinned: bit-identical to the unmodified reference's HCopy (oracle/_ref/bin/HCopy) on
tests/golden/qualifiers_*.npz (tests/golden/make_qualifier_golden.py; tests/test_oracle_golden.py).

`decompress` restates the loader's treatment of `_C` compressed files (HParm.c:3489-3494) and `crc` the `_K` check sum
(UpdateCRCC, HParm.c:3357-3380).  Pinned: bit-identical to what the unmodified reference's HCopy reads back from the
compressed files it wrote itself (tests/golden/compressed_*.npz, tests/golden/make_compressed_golden.py).
"""
import numpy as np

F = np.float32


def decompress(shorts: np.ndarray, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """HParm.c:3492-3493: v[j] = ((float)s[j] + cf->B[j]) / cf->A[j], two float operations, each rounded."""
    return ((shorts.astype(F) + np.asarray(B, F)).astype(F) / np.asarray(A, F)).astype(F)


def crc(words) -> int:
    """HParm.c:3357-3380 over the file's 16-bit words after the header, in file order, starting from 0."""
    c = 0
    for w in np.asarray(words, dtype=np.uint64).tolist():
        c = (c * 65536 + w) % 36897                       # :3368-3373
    return c


def regress(src: np.ndarray, win: int, simple: bool) -> np.ndarray:
    """HSigP.c:827-857 with head = tail = 0 blocks joined: first / last row replicated."""
    T = src.shape[0]
    idx = np.arange(T)
    sigma = F(0)
    for th in range(1, win + 1):
        sigma = F(sigma + F(th * th))                     # :835-836
    sigma = F(sigma * F(2))                               # :837
    acc = np.zeros_like(src, dtype=F)
    fw = bk = src
    for th in range(1, win + 1):
        fw = src[np.minimum(idx + th, T - 1)]             # :844-845: the pointers stop at the ends
        bk = src[np.maximum(idx - th, 0)]
        if not simple:
            acc = (acc + (F(th) * (fw - bk).astype(F)).astype(F)).astype(F)   # :846
    if simple:
        return ((fw - bk).astype(F) / F(2 * win)).astype(F)                   # :849
    return (acc / sigma).astype(F)                                            # :851


def expand(static: np.ndarray, del_win=0, acc_win=0, third_win=0, simple=False, zero_mean_cols=0,
           suppress_energy=False) -> np.ndarray:
    """AddQualifiers: deltas over all static columns, accelerations over the deltas, thirds over the
    accelerations (HParm.c:1664-1697), then FZeroMean over the leading columns (:1707-1722); with _N the last static
    column (energy / c0) is left out of the observation that is delivered (HParm.c:2882, :4655-4656)."""
    x = np.ascontiguousarray(static, dtype=F)
    cols = [x]
    for w in (del_win, acc_win, third_win):
        if w <= 0:
            break
        cols.append(regress(cols[-1], w, simple))
    out = np.concatenate(cols, axis=1)
    for c in range(zero_mean_cols):
        s = 0.0
        for v in out[:, c]:                               # HSigP.c:811-815: double sum, in order
            s += float(v)
        mean = F(s / float(out.shape[0]))
        out[:, c] = (out[:, c] - mean).astype(F)
    if suppress_energy:
        out = np.delete(out, x.shape[1] - 1, axis=1)
    return out
