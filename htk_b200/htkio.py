"""HTK on-disk formats needed either side of the E-step: parameter files, MLFs,
text MMFs, HMM lists and the ``HER$.acc`` accumulator dump.

These are format restatements written from the reference's readers/writers:

* parameter files -- 12-byte big-endian header + big-endian float32 rows
  (HTKLib/HWave.c:1399-1433, HTKLib/HParm.c:4616 ``ReadAsTable``);
* text MMF -- ``GetHMMDef``/``GetStream``/``GetMixture``/``GetTransMat``
  (HTKLib/HModel.c:2060, :1850, :1771, :2003) and ``SaveHMMSet`` (:4979);
* accumulator dump -- ``DumpAccs``/``LoadAccs`` (HTKLib/HTrain.c:1454-1505,
  :1626-1687) plus HERest's trailer (HTKTools/HERest.c:544-549);
* physical-HMM scan order -- ``NewHMMScan``/``GoNextHMM`` over the macro hash
  table (HTKLib/HUtil.c:246-296, HTKLib/HModel.c:3314-3321, :3384).

HERest keeps owning these formats in production; the Python
versions exist for tests, fixtures, bench.py and the ``-p`` dump bridge.
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

LZERO = -1.0e10
LSMALL = -0.5e10
MINLARG = 2.45e-308
MINMIX = 1.0e-5
LMINMIX = -11.5129254649702
TPI = 6.28318530717959
MACHASHSIZE = 250007

UPMEANS, UPVARS, UPTRANS, UPMIXES = 1, 2, 4, 8

# parmKind base codes and qualifier bits (HTKLib/HParm.h)
_BASE = {"WAVEFORM": 0, "LPC": 1, "LPREFC": 2, "LPCEPSTRA": 3, "LPDELCEP": 4, "IREFC": 5,
         "MFCC": 6, "FBANK": 7, "MELSPEC": 8, "USER": 9, "DISCRETE": 10, "PLP": 11}
_QUAL = {"E": 0o100, "N": 0o200, "D": 0o400, "A": 0o1000, "C": 0o2000, "Z": 0o4000,
         "K": 0o10000, "0": 0o20000, "V": 0o40000, "T": 0o100000}


def parm_kind_code(name: str) -> int:
    parts = name.split("_")
    code = _BASE[parts[0]]
    for q in parts[1:]:
        code |= _QUAL[q]
    return code


# --------------------------------------------------------------------------- features

def write_htk_features(path: str, feat: np.ndarray, parm_kind: str = "MFCC_0_D_A",
                       samp_period: int = 100000, with_crc: bool = False) -> None:
    """with_crc: the `_K` form HCopy writes by default (SAVEWITHCRC = T): kind | HASCRCC and a 16-bit check sum over the
    payload's 16-bit words in file order (UpdateCRCC, HParm.c:3357-3380) after the last row."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    T, D = feat.shape
    body = feat.astype(">f4").tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack(">iihH", T, samp_period, D * 4, parm_kind_code(parm_kind) | (_QUAL["K"] if with_crc else 0)))
        f.write(body)
        if with_crc:
            f.write(struct.pack(">H", htk_crc(np.frombuffer(body, dtype=">u2"))))


def read_htk_features(path: str) -> Tuple[np.ndarray, int, int]:
    """Returns (feat[T, D] float32, sampPeriod, parmKind).  Handles the _K CRC trailer;
    does NOT expand qualifiers (HParm stays the owner of that)."""
    with open(path, "rb") as f:
        raw = f.read()
    T, period, size, kind = struct.unpack(">iihh", raw[:12])
    kind &= 0xFFFF
    D = size // 4
    feat = np.frombuffer(raw, dtype=">f4", count=T * D, offset=12).astype(np.float32).reshape(T, D)
    return feat, period, kind


# ---- `_C` compressed parameter files (HASCOMPX) and the `_K` check sum (HASCRCC) -----------------------------------
# Layout (HTKLib/HParm.c SaveBuffer; read side OpenParmChannel :3680-3699): 12-byte header whose nSamples counts 4 extra
# rows, sampSize = 2 * cols; then vector A[cols], vector B[cols] (floats), then T rows of cols 16-bit integers; with _K a
# 16-bit check sum over every 16-bit word after the header, in file order (UpdateCRCC, :3357-3380).  All big-endian.

def htk_crc(words_be: np.ndarray, crc: int = 0) -> int:
    """UpdateCRCC (HParm.c:3357-3380) over 16-bit words taken in FILE order (`>u2` view of the payload)."""
    for w in np.asarray(words_be, dtype=np.uint64).tolist():
        crc = (crc * 65536 + w) % 36897
    return crc


def compress_params(feat: np.ndarray):
    """CalcCompress + CompressPBlock (HParm.c:4892-4960) in the reference's arithmetic: float min / max and differences,
    the quotients in double rounded to float, `x = f * A - B` in float, rounding half away from zero.
    Returns (shorts[T, cols] int16, A[cols], B[cols])."""
    f = np.ascontiguousarray(feat, dtype=np.float32)
    mx, mn = f.max(axis=0), f.min(axis=0)
    rng = (mx - mn).astype(np.float32)
    flat = rng == 0
    safe = np.where(flat, np.float32(1), rng).astype(np.float64)
    A = np.where(flat, 1.0, 2.0 * 32767.0 / safe).astype(np.float32)
    B = np.where(flat, mx.astype(np.float64), (mx + mn).astype(np.float32).astype(np.float64) * 32767.0 / safe).astype(np.float32)
    x = (f * A).astype(np.float32) - B                      # float products and differences, each rounded (no FMA in the reference build)
    x = x.astype(np.float32).astype(np.float64)
    ix = np.where(x < 0.0, x - 0.5, x + 0.5).astype(np.int64)   # C's (int) truncates toward zero
    if ix.min() < -32767 or ix.max() > 32767:
        raise ValueError("CompressPBlock: short out of range (HError 6393)")
    return ix.astype(np.int16), A, B


def decompress_params(shorts: np.ndarray, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """What the reference's loader makes of a compressed row: v[j] = ((float)s[j] + B[j]) / A[j] (HParm.c:3489-3494)."""
    return ((shorts.astype(np.float32) + B.astype(np.float32)).astype(np.float32) / A.astype(np.float32)).astype(np.float32)


def write_htk_compressed(path: str, feat: np.ndarray, parm_kind: str = "MFCC_0_D_A", samp_period: int = 100000,
                         with_crc: bool = True) -> None:
    """`feat` saved as HCopy with SAVECOMPRESSED = T (and SAVEWITHCRC) saves it; parm_kind WITHOUT _C / _K."""
    s, A, B = compress_params(feat)
    T, D = s.shape
    kind = parm_kind_code(parm_kind) | _QUAL["C"] | (_QUAL["K"] if with_crc else 0)
    body = A.astype(">f4").tobytes() + B.astype(">f4").tobytes() + s.astype(">i2").tobytes()
    with open(path, "wb") as f:
        f.write(struct.pack(">iihH", T + 4, samp_period, D * 2, kind))
        f.write(body)
        if with_crc:
            f.write(struct.pack(">H", htk_crc(np.frombuffer(body, dtype=">u2"))))


def read_htk_compressed(path: str):
    """Returns (shorts[T, cols] int16, A, B, sampPeriod, parmKind, crc_ok or None) of a `_C` file, nothing decoded."""
    with open(path, "rb") as f:
        raw = f.read()
    n, period, size, kind = struct.unpack(">iihH", raw[:12])
    if not kind & _QUAL["C"]:
        raise ValueError("%s is not a compressed parameter file" % path)
    D, T = size // 2, n - 4
    A = np.frombuffer(raw, dtype=">f4", count=D, offset=12).astype(np.float32)
    B = np.frombuffer(raw, dtype=">f4", count=D, offset=12 + 4 * D).astype(np.float32)
    s = np.frombuffer(raw, dtype=">i2", count=T * D, offset=12 + 8 * D).astype(np.int16).reshape(T, D)
    ok = None
    if kind & _QUAL["K"]:
        end = 12 + 8 * D + 2 * T * D
        ok = htk_crc(np.frombuffer(raw[12:end], dtype=">u2")) == struct.unpack(">H", raw[end:end + 2])[0]
    return s, A, B, period, kind, ok


def write_mlf(path: str, labels: Dict[str, Sequence[str]]) -> None:
    with open(path, "w") as f:
        f.write("#!MLF!#\n")
        for utt, labs in labels.items():
            f.write('"*/%s.lab"\n' % utt)
            for l in labs:
                f.write(l + "\n")
            f.write(".\n")


def read_label_file(path: str) -> List[str]:
    """One label file: optional start/end times, then the label name."""
    out = []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if len(p) >= 3 and p[0].lstrip("-").isdigit() and p[1].lstrip("-").isdigit():
                out.append(p[2])
            else:
                out.append(p[0])
    return out


# --------------------------------------------------------------------------- model set

@dataclass
class Gaussian:
    mean: np.ndarray                 # [D] float32
    var: np.ndarray                  # [D] float32 VARIANCES (as in an MMF)
    gconst: Optional[float] = None   # value given in the MMF, if any
    mean_key: Optional[str] = None   # ~u macro name when shared
    var_key: Optional[str] = None    # ~v macro name when shared


@dataclass
class State:
    mixes: List[Tuple[float, Gaussian]]   # (weight, pdf)
    name: Optional[str] = None            # ~s macro name when shared


@dataclass
class TransMat:
    prob: np.ndarray                      # [N][N] probabilities
    name: Optional[str] = None            # ~t macro name when shared


@dataclass
class HMM:
    name: str
    states: List[State]                   # emitting states 2..N-1
    trans: TransMat


@dataclass
class HMMSetDef:
    vec_size: int
    parm_kind: str = "MFCC_0_D_A"
    hmms: List[HMM] = field(default_factory=list)                 # physical, in LIST order
    logical: List[Tuple[str, str]] = field(default_factory=list)  # (logical, physical)

    def physical_names(self) -> List[str]:
        return [h.name for h in self.hmms]

    def logical_map(self) -> Dict[str, str]:
        m = {h.name: h.name for h in self.hmms}
        m.update(dict(self.logical))
        return m


def htk_hash(name: str) -> int:
    """HModel.c:3314-3321 (char is signed on x86-64, arithmetic is unsigned 32-bit)."""
    h = 0
    for ch in name.encode("latin-1"):
        c = ch - 256 if ch > 127 else ch
        h = (c + 31 * h) & 0xFFFFFFFF
    return h % MACHASHSIZE


def scan_order(names_in_creation_order: Sequence[str]) -> List[str]:
    """Order in which NewHMMScan/GoNextHMM visit physical HMMs: buckets ascending,
    newest first inside a bucket (NewMacro pushes at the head, HModel.c:3384)."""
    keyed = [(htk_hash(n), -i, n) for i, n in enumerate(names_in_creation_order)]
    keyed.sort()
    return [n for _, _, n in keyed]


# --------------------------------------------------------------------------- MMF text

def _fmt_vec(v) -> str:
    return " ".join("%.6e" % float(x) for x in v)


def write_mmf(path: str, hs: HMMSetDef, with_gconst: bool = False) -> None:
    """Text MMF with ~s / ~t macros for shared structures (fixture format of SURVEY.md 8d)."""
    D = hs.vec_size
    seen_s, seen_t = set(), set()
    with open(path, "w") as f:
        f.write("~o\n<STREAMINFO> 1 %d\n<VECSIZE> %d<NULLD><%s><DIAGC>\n" % (D, D, hs.parm_kind))

        def emit_state(st: State):
            M = len(st.mixes)
            if M > 1:
                f.write("<NUMMIXES> %d\n" % M)
            for m, (w, g) in enumerate(st.mixes, 1):
                if M > 1:
                    f.write("<MIXTURE> %d %.6e\n" % (m, w))
                f.write("<MEAN> %d\n %s\n<VARIANCE> %d\n %s\n" % (D, _fmt_vec(g.mean), D, _fmt_vec(g.var)))
                if with_gconst and g.gconst is not None:
                    f.write("<GCONST> %.6e\n" % g.gconst)

        def emit_trans(tm: TransMat):
            N = tm.prob.shape[0]
            f.write("<TRANSP> %d\n" % N)
            for i in range(N):
                f.write(" " + _fmt_vec(tm.prob[i]) + "\n")

        for h in hs.hmms:
            if h.trans.name and h.trans.name not in seen_t:
                seen_t.add(h.trans.name)
                f.write('~t "%s"\n' % h.trans.name)
                emit_trans(h.trans)
            for st in h.states:
                if st.name and st.name not in seen_s:
                    seen_s.add(st.name)
                    f.write('~s "%s"\n' % st.name)
                    emit_state(st)
        for h in hs.hmms:
            f.write('~h "%s"\n<BEGINHMM>\n<NUMSTATES> %d\n' % (h.name, len(h.states) + 2))
            for j, st in enumerate(h.states, 2):
                f.write("<STATE> %d\n" % j)
                if st.name:
                    f.write('~s "%s"\n' % st.name)
                else:
                    emit_state(st)
            if h.trans.name:
                f.write('~t "%s"\n' % h.trans.name)
            else:
                emit_trans(h.trans)
            f.write("<ENDHMM>\n")


def write_hmm_list(path: str, hs: HMMSetDef) -> None:
    with open(path, "w") as f:
        for h in hs.hmms:
            f.write(h.name + "\n")
        for lg, ph in hs.logical:
            f.write("%s %s\n" % (lg, ph))


_TOK = re.compile(r'~[a-z]|<[A-Za-z0-9_]+>|"[^"]*"|[^\s<>"~]+')


class _Toks:
    def __init__(self, text: str):
        self.t = _TOK.findall(text)
        self.i = 0

    def peek(self) -> Optional[str]:
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self) -> str:
        v = self.t[self.i]
        self.i += 1
        return v

    def floats(self, n: int) -> np.ndarray:
        v = np.array([float(x) for x in self.t[self.i:self.i + n]], dtype=np.float64)
        self.i += n
        return v


_KINDTOK = re.compile(r"<(MFCC|USER|PLP|FBANK|LPC|LPCEPSTRA|MELSPEC|LPREFC|LPDELCEP|IREFC)[A-Z0-9_]*>")


def read_mmf(paths: Sequence[str], hmm_list: Optional[Sequence[str]] = None) -> HMMSetDef:
    """Parse text MMFs / single-HMM definition files (diagonal covariance, one stream).

    ``paths`` may hold several files (HERest -H a -H b, or one file per model as HInit
    writes them).  Physical order = ``hmm_list`` order when given, else file order.
    """
    states: Dict[str, State] = {}
    trans: Dict[str, TransMat] = {}
    means: Dict[str, np.ndarray] = {}
    varis: Dict[str, np.ndarray] = {}
    hmms: Dict[str, HMM] = {}
    D = None
    kind = "USER"

    def parse_mean(tk):
        nxt = tk.peek()
        if nxt == "~u":
            tk.next()
            return means[tk.next().strip('"')], None
        assert tk.next().upper() == "<MEAN>"
        n = int(tk.next())
        return tk.floats(n).astype(np.float32), None

    def parse_var(tk):
        if tk.peek() == "~v":
            tk.next()
            key = tk.next().strip('"')
            return varis[key], key
        assert tk.next().upper() == "<VARIANCE>"
        n = int(tk.next())
        return tk.floats(n).astype(np.float32), None

    def parse_mixpdf(tk) -> Gaussian:
        if tk.peek() and tk.peek().upper() == "<RCLASS>":
            tk.next(); tk.next()
        mkey = None
        if tk.peek() == "~u":
            tk.next()
            mkey = tk.next().strip('"')
            mean = means[mkey]
        else:
            mean, _ = parse_mean(tk)
        var, vkey = parse_var(tk)
        g = Gaussian(mean, var, None, mkey, vkey)
        if tk.peek() and tk.peek().upper() == "<GCONST>":
            tk.next()
            g.gconst = float(tk.next())
        return g

    def parse_state(tk) -> State:
        M = 1
        while tk.peek() and tk.peek().upper() in ("<NUMMIXES>", "<SWEIGHTS>", "<STREAM>"):
            t = tk.next().upper()
            if t == "<NUMMIXES>":
                M = int(tk.next())
            elif t == "<SWEIGHTS>":
                n = int(tk.next()); tk.floats(n)
            else:
                tk.next()
        if M == 1 and (tk.peek() or "").upper() != "<MIXTURE>":
            return State([(1.0, parse_mixpdf(tk))])
        mixes: List[Optional[Tuple[float, Gaussian]]] = [None] * M
        while tk.peek() and tk.peek().upper() == "<MIXTURE>":
            tk.next()
            m = int(tk.next()); w = float(tk.next())
            mixes[m - 1] = (w, parse_mixpdf(tk))
        if any(x is None for x in mixes):
            raise ValueError("MMF: state with missing mixture components is not supported")
        return State(mixes)  # type: ignore[arg-type]

    def parse_trans(tk) -> TransMat:
        assert tk.next().upper() == "<TRANSP>"
        N = int(tk.next())
        return TransMat(tk.floats(N * N).reshape(N, N))

    order: List[str] = []
    for p in paths:
        with open(p) as f:
            text = f.read()
        mk = _KINDTOK.search(text)
        if mk:
            kind = mk.group(0)[1:-1]
        tk = _Toks(text)
        while tk.peek() is not None:
            t = tk.next()
            tu = t.upper()
            if t == "~o":
                while tk.peek() is not None and not tk.peek().startswith("~"):
                    x = tk.next().upper()
                    if x == "<VECSIZE>":
                        D = int(tk.next())
                    elif x == "<STREAMINFO>":
                        n = int(tk.next())
                        if n != 1:
                            raise ValueError("only single-stream sets are on the path")
                        tk.next()
            elif t == "~s":
                name = tk.next().strip('"')
                st = parse_state(tk); st.name = name
                states[name] = st
            elif t == "~t":
                name = tk.next().strip('"')
                tm = parse_trans(tk); tm.name = name
                trans[name] = tm
            elif t == "~u":
                name = tk.next().strip('"')
                means[name], _ = parse_mean(tk)
            elif t == "~v":
                name = tk.next().strip('"')
                assert tk.next().upper() == "<VARIANCE>"
                n = int(tk.next())
                varis[name] = tk.floats(n).astype(np.float32)
            elif t == "~h" or tu == "<BEGINHMM>":
                if t == "~h":
                    name = tk.next().strip('"')
                    assert tk.next().upper() == "<BEGINHMM>"
                else:
                    import os
                    name = os.path.basename(p)
                sts: List[State] = []
                tm = None
                N = None
                while True:
                    x = tk.next()
                    xu = x.upper()
                    if xu == "<NUMSTATES>":
                        N = int(tk.next())
                    elif xu == "<STATE>":
                        tk.next()
                        if tk.peek() == "~s":
                            tk.next()
                            sts.append(states[tk.next().strip('"')])
                        else:
                            sts.append(parse_state(tk))
                    elif x == "~t":
                        tm = trans[tk.next().strip('"')]
                    elif xu == "<TRANSP>":
                        tk.i -= 1
                        tm = parse_trans(tk)
                    elif xu == "<ENDHMM>":
                        break
                    elif xu in ("<VECSIZE>",):
                        D = int(tk.next())
                    # other global options inside the definition are ignored
                assert N is not None and len(sts) == N - 2 and tm is not None
                hmms[name] = HMM(name, sts, tm)
                order.append(name)
            # anything else (e.g. option tokens) is skipped
    if D is None:
        D = len(next(iter(hmms.values())).states[0].mixes[0][1].mean)
    hs = HMMSetDef(D, kind)
    if hmm_list is None:
        hs.hmms = [hmms[n] for n in order]
    else:
        phys: List[str] = []
        for line in hmm_list:
            p = line.split()
            if not p:
                continue
            lg, ph = p[0], (p[1] if len(p) > 1 else p[0])
            if ph not in phys:
                phys.append(ph)
            if lg != ph:
                hs.logical.append((lg, ph))
        hs.hmms = [hmms[n] for n in phys]
    return hs


# --------------------------------------------------------------------------- acc dumps

def _acc_records(hs: HMMSetDef, uflags: int, order: Sequence[str]):
    """Yield, in dump order, ('name', hmm) then ('wt', state) / ('mu', gauss) / ('va', gauss)
    / ('tr', transmat) / ('mark',) following DumpAccs' seen-flag logic."""
    byname = {h.name: h for h in hs.hmms}
    seen_state, seen_mean, seen_var, seen_tr, seen_pdf = set(), set(), set(), set(), set()
    for n in order:
        h = byname[n]
        yield ("name", h)
        for st in h.states:
            if id(st) in seen_state:
                continue
            seen_state.add(id(st))
            yield ("wt", st)
            for _, g in st.mixes:
                if id(g) in seen_pdf:
                    continue
                seen_pdf.add(id(g))
                if (uflags & UPMEANS) and id(g.mean) not in seen_mean:
                    seen_mean.add(id(g.mean))
                    yield ("mu", g)
                if (uflags & UPVARS) and id(g.var) not in seen_var:
                    seen_var.add(id(g.var))
                    yield ("va", g)
        if id(h.trans) not in seen_tr:
            seen_tr.add(id(h.trans))
            yield ("tr", h.trans)
        yield ("mark",)


def read_acc_dump(path: str, hs: HMMSetDef, flat, uflags: int = 15):
    """Decode a binary HER$.acc into a flat float64 array in ``flat.layout`` order.
    Returns (acc, totalPr, totalT)."""
    with open(path, "rb") as f:
        raw = f.read()
    L = flat.layout
    acc = np.zeros(L.count, dtype=np.float64)
    pos = 0
    D = hs.vec_size

    def rd_f(n):
        nonlocal pos
        v = np.frombuffer(raw, dtype=">f4", count=n, offset=pos).astype(np.float64)
        pos += 4 * n
        return v

    def rd_i():
        nonlocal pos
        v = struct.unpack(">i", raw[pos:pos + 4])[0]
        pos += 4
        return v

    # physical order is read from the file itself (names are in the records)
    order = []
    p = 0
    names = set(h.name for h in hs.hmms)
    for m in re.finditer(rb'"([^"\n]+)"\n', raw):
        nm = m.group(1).decode("latin-1")
        if nm in names and nm not in order:
            order.append(nm)
    cur = None
    for rec in _acc_records(hs, uflags, order):
        kind = rec[0]
        if kind == "name":
            cur = rec[1]
            tag = ('"%s"\n' % cur.name).encode("latin-1")
            if raw[pos:pos + len(tag)] != tag:
                # names with special characters are written escaped; fall back to a scan
                nl = raw.index(b"\n", pos)
                pos = nl + 1
            else:
                pos += len(tag)
            acc[L.numEgs + flat.hmm_index[cur.name]] = rd_i()
        elif kind == "wt":
            st = rec[1]
            s = flat.state_index[id(st)]
            M = len(st.mixes)
            acc[L.wtC + flat.stateMixOff[s]: L.wtC + flat.stateMixOff[s] + M] = rd_f(M)
            acc[L.wtOcc + s] = rd_f(1)[0]
        elif kind == "mu":
            i = flat.mean_index[id(rec[1].mean)]
            acc[L.muSum + i * D: L.muSum + (i + 1) * D] = rd_f(D)
            acc[L.muOcc + i] = rd_f(1)[0]
        elif kind == "va":
            i = flat.var_index[id(rec[1].var)]
            acc[L.vaSum + i * D: L.vaSum + (i + 1) * D] = rd_f(D)
            acc[L.vaOcc + i] = rd_f(1)[0]
        elif kind == "tr":
            tid = flat.trans_index[id(rec[1])]
            N = int(flat.transN[tid])
            o = L.tran + int(flat.tranAccOff[tid])
            acc[o:o + N * N] = rd_f(N * N)
            o = L.tranOcc + int(flat.tranOccOff[tid])
            acc[o:o + N] = rd_f(N)
        else:
            mark = rd_i()
            if mark != 123456:
                raise ValueError("acc dump: bad marker %d in %s" % (mark, cur.name if cur else "?"))
    total_pr = float(rd_f(1)[0])
    total_t = rd_i()
    acc[L.totalPr] = total_pr
    acc[L.totalT] = total_t
    return acc, total_pr, total_t


def read_acc_dump_flat(path: str, flat, uflags: int = 15):
    """Decode a binary HER$.acc (HTrain.c:1454-1505 DumpAccs + HERest.c:544-549 trailer) straight into a
    flat float64 array in ``flat.layout`` order, for a FlatModel without an HMMSetDef behind it (the
    BASELINE-size synthetic sets of synth.make_flat_*: no Python object per Gaussian).  The physical order
    is the one of the file (names are in the records); sharing follows DumpAccs' seen flags: a tied state,
    a mean vector, a variance vector and a transition matrix appear once, inside the first HMM that uses
    them.  Returns (acc, totalPr, totalT)."""
    with open(path, "rb") as f:
        raw = f.read()
    L = flat.layout
    D = flat.D
    acc = np.zeros(L.count, dtype=np.float64)
    index = {n: i for i, n in enumerate(flat.names)}
    seen_state = np.zeros(flat.J, bool); seen_tr = np.zeros(flat.numTrans, bool)
    seen_mu = np.zeros(flat.numMeanAcc, bool); seen_va = np.zeros(flat.numVarAcc, bool)
    seen_g = np.zeros(flat.G, bool)
    pos = 0
    n_hmm = 0
    while n_hmm < flat.P:
        nl = raw.index(b"\n", pos)
        name = raw[pos:nl].decode("latin-1")
        if name.startswith('"') and name.endswith('"'):
            name = name[1:-1]
        p_ = index[name.replace("\\", "")]
        pos = nl + 1
        acc[L.numEgs + p_] = struct.unpack(">i", raw[pos:pos + 4])[0]
        pos += 4
        for s in flat.hmmState[flat.hmmStateOff[p_]:flat.hmmStateOff[p_ + 1]]:
            s = int(s)
            if seen_state[s]:
                continue
            seen_state[s] = True
            o, e = int(flat.stateMixOff[s]), int(flat.stateMixOff[s + 1])
            M = e - o
            v = np.frombuffer(raw, dtype=">f4", count=M + 1, offset=pos); pos += 4 * (M + 1)
            acc[L.wtC + o:L.wtC + e] = v[:M]; acc[L.wtOcc + s] = v[M]
            for g in flat.mixGauss[o:e]:
                g = int(g)
                if seen_g[g]:
                    continue
                seen_g[g] = True
                i = int(flat.meanId[g])
                if (uflags & UPMEANS) and not seen_mu[i]:
                    seen_mu[i] = True
                    v = np.frombuffer(raw, dtype=">f4", count=D + 1, offset=pos); pos += 4 * (D + 1)
                    acc[L.muSum + i * D:L.muSum + (i + 1) * D] = v[:D]; acc[L.muOcc + i] = v[D]
                i = int(flat.varId[g])
                if (uflags & UPVARS) and not seen_va[i]:
                    seen_va[i] = True
                    v = np.frombuffer(raw, dtype=">f4", count=D + 1, offset=pos); pos += 4 * (D + 1)
                    acc[L.vaSum + i * D:L.vaSum + (i + 1) * D] = v[:D]; acc[L.vaOcc + i] = v[D]
        t = int(flat.hmmTrans[p_])
        if not seen_tr[t]:
            seen_tr[t] = True
            N = int(flat.transN[t])
            v = np.frombuffer(raw, dtype=">f4", count=N * N + N, offset=pos); pos += 4 * (N * N + N)
            oa = L.tran + int(flat.tranAccOff[t]); ob = L.tranOcc + int(flat.tranOccOff[t])
            acc[oa:oa + N * N] = v[:N * N]; acc[ob:ob + N] = v[N * N:]
        mark = struct.unpack(">i", raw[pos:pos + 4])[0]; pos += 4
        if mark != 123456:
            raise ValueError("acc dump: bad marker %d after %s" % (mark, name))
        n_hmm += 1
    total_pr = float(np.frombuffer(raw, dtype=">f4", count=1, offset=pos)[0]); pos += 4
    total_t = struct.unpack(">i", raw[pos:pos + 4])[0]
    acc[L.totalPr] = total_pr; acc[L.totalT] = total_t
    return acc, total_pr, total_t


def write_acc_dump(path: str, hs: HMMSetDef, flat, acc: np.ndarray, uflags: int = 15,
                   order: Optional[Sequence[str]] = None) -> None:
    """Write flat FP64 accumulators as a binary HER$.acc that stock ``HERest -p 0`` loads
    (HTrain.c:1454-1505 + HERest.c:544-549).  ``order`` defaults to HTK's scan order for
    a set whose HMM list is ``hs.hmms`` order."""
    L = flat.layout
    D = hs.vec_size
    if order is None:
        # MakeHMMSet creates macros in list order: for each line the physical 'h' macro
        # (if new) then the logical 'l' macro; only 'h' macros matter for the scan.
        order = scan_order(hs.physical_names())
    out = bytearray()

    def wr_f(v):
        out.extend(np.asarray(v, dtype=np.float64).astype(">f4").tobytes())

    def wr_i(v):
        out.extend(struct.pack(">i", int(v)))

    for rec in _acc_records(hs, uflags, order):
        kind = rec[0]
        if kind == "name":
            h = rec[1]
            out.extend(('"%s"\n' % h.name).encode("latin-1"))
            wr_i(round(acc[L.numEgs + flat.hmm_index[h.name]]))
        elif kind == "wt":
            st = rec[1]
            s = flat.state_index[id(st)]
            M = len(st.mixes)
            wr_f(acc[L.wtC + flat.stateMixOff[s]: L.wtC + flat.stateMixOff[s] + M])
            wr_f([acc[L.wtOcc + s]])
        elif kind == "mu":
            i = flat.mean_index[id(rec[1].mean)]
            wr_f(acc[L.muSum + i * D: L.muSum + (i + 1) * D]); wr_f([acc[L.muOcc + i]])
        elif kind == "va":
            i = flat.var_index[id(rec[1].var)]
            wr_f(acc[L.vaSum + i * D: L.vaSum + (i + 1) * D]); wr_f([acc[L.vaOcc + i]])
        elif kind == "tr":
            tid = flat.trans_index[id(rec[1])]
            N = int(flat.transN[tid])
            o = L.tran + int(flat.tranAccOff[tid]); wr_f(acc[o:o + N * N])
            o = L.tranOcc + int(flat.tranOccOff[tid]); wr_f(acc[o:o + N])
        else:
            wr_i(123456)
    wr_f([acc[L.totalPr]])
    wr_i(round(acc[L.totalT]))
    with open(path, "wb") as f:
        f.write(bytes(out))
