"""ctypes front-end of the CPU oracle (oracle/mcaller_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates, on the CPU, what the reference's `extract_features`
(extract_contexts.py:110-304) and `aggregate_by_pos` (make_bed.py:67-164) compute, and renders
the same text rows, so outputs can be compared byte for byte with the golden vectors.

Independent of the product package on purpose: it has its own FASTA/FASTQ readers, reference
marking, model export and row formatting.
"""
import ctypes as C
import os
import pickle
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmcoracle.so")
MAXK = 16


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "mcaller_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


class _Contig(C.Structure):
    _fields_ = [("name", C.c_char_p), ("fwd", C.c_char_p), ("rev", C.c_char_p), ("len", C.c_int64)]


class _Qual(C.Structure):
    _fields_ = [("key", C.c_char_p), ("qual", C.c_double)]


class _Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_in", C.c_int32), ("n_layers", C.c_int32), ("sizes", C.c_int32 * 8),
                ("hidden_act", C.c_int32), ("weights", C.c_void_p), ("biases", C.c_void_p),
                ("n_trees", C.c_int32), ("tree_off", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p),
                ("feature", C.c_void_p), ("threshold", C.c_void_p), ("leaf_p1", C.c_void_p)]


class _Call(C.Structure):
    _fields_ = [("kind", C.c_int32), ("chrom_cid", C.c_int32), ("win_cid", C.c_int32), ("rev", C.c_int32),
                ("read_off", C.c_int64), ("read_len", C.c_int32), ("mpos", C.c_int32), ("n_empty", C.c_int32),
                ("empty_mask", C.c_uint32), ("model_sel", C.c_int32), ("label", C.c_int32),
                ("feat", C.c_double * (MAXK + 1)), ("prob", C.c_double), ("context", C.c_char * (2 * MAXK))]


class _Locus(C.Structure):
    _fields_ = [("key_off", C.c_int64 * 5), ("key_len", C.c_int32 * 5), ("depth", C.c_int64), ("meth", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_extract.restype = C.c_int64
        _lib.orc_aggregate.restype = C.c_int64
        _lib.orc_predict.restype = C.c_double
        _lib.orc_np_mean.restype = C.c_double
        _lib.orc_np_round.restype = C.c_double
        _lib.orc_np_round.argtypes = [C.c_double, C.c_int]
        assert _lib.orc_sizeof_call() == C.sizeof(_Call), (_lib.orc_sizeof_call(), C.sizeof(_Call))
        assert _lib.orc_sizeof_locus() == C.sizeof(_Locus)
    return _lib


# ---- inputs -------------------------------------------------------------------------------------

def read_fasta(path):
    seqs, name, parts = {}, None, []
    for ln in open(path):
        if ln[:1] == ">":
            if name is not None:
                seqs[name] = "".join(parts).upper()
            t = ln[1:].split()
            name, parts = (t[0] if t else ""), []
        elif name is not None:
            parts.append(ln.strip())
    if name is not None:
        seqs[name] = "".join(parts).upper()
    return seqs


def read_fastq_quals(path):
    """read_qual.py:6-19: {id.split(':')[0].split('_')[0]: mean phred}."""
    import gzip
    op = gzip.open if path.find(".gz") != -1 else open
    out = {}
    with op(path, "rt") as fh:
        while True:
            h = fh.readline()
            if not h:
                break
            if not h.strip():
                continue
            fh.readline()
            fh.readline()
            q = fh.readline().rstrip("\n")
            toks = h[1:].split()
            rid = (toks[0] if toks else "").split(":")[0].split("_")[0]
            out[rid] = float(np.mean([ord(c) - 33 for c in q]))
    return out


_CB = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "M": "M"}


def _rc(s):
    return "".join(_CB[c] for c in reversed(s))


def mark(seq, base, motif=None, positions=None, contig=None):
    """extract_contexts.py:33-73 in plain Python string operations."""
    if not positions and motif:
        f = seq.replace(motif, "M".join(motif.split(base)))
        rm = _rc(motif)
        r = seq.replace(rm, "M".join(rm.split(_CB[base])))
        return f, r
    rows = [x.split() for x in open(positions).read().split("\n")]
    outs = []
    for st, b in (("+", base), ("-", _CB[base])):
        buf = list(seq)
        for x in rows:
            if len(x) > 1 and x[2] == st and x[0] == contig:
                p = int(x[1])
                if buf[p] != b and buf[p] != "M":
                    raise ValueError("Base %d does not correspond to methylated base" % p)
                buf[p] = "M"
        outs.append("".join(buf))
    return outs[0], outs[1]


class _Aliasing(pickle.Unpickler):
    _MAP = {"sklearn.neural_network.multilayer_perceptron": "sklearn.neural_network._multilayer_perceptron",
            "sklearn.preprocessing.label": "sklearn.preprocessing._label",
            "sklearn.ensemble.forest": "sklearn.ensemble._forest", "sklearn.tree.tree": "sklearn.tree._classes",
            "sklearn.linear_model.logistic": "sklearn.linear_model._logistic"}

    def find_class(self, module, name):
        return super().find_class(self._MAP.get(module, module), name)


def load_pickle(path):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(path, "rb") as fh:
            return _Aliasing(fh, encoding="latin").load()


_ACT = {"identity": 0, "logistic": 1, "tanh": 2, "relu": 3}


def export_model(est):
    """sklearn estimator -> (_Model, keepalive list)."""
    m = _Model()
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    tn = type(est).__name__
    if tn == "MLPClassifier":
        m.kind = 0
        sizes = [est.coefs_[0].shape[0]] + [c.shape[1] for c in est.coefs_]
        assert sizes[-1] == 1 and est.out_activation_ == "logistic" and len(sizes) <= 8
        m.n_in, m.n_layers = sizes[0], len(est.coefs_)
        for i, s in enumerate(sizes):
            m.sizes[i] = s
        m.hidden_act = _ACT[est.activation]
        m.weights = arr(np.concatenate([c.ravel() for c in est.coefs_]), np.float64)
        m.biases = arr(np.concatenate([b.ravel() for b in est.intercepts_]), np.float64)
    elif tn == "LogisticRegression":
        m.kind, m.n_in = 1, est.coef_.shape[1]
        m.weights, m.biases = arr(est.coef_.ravel(), np.float64), arr(est.intercept_.ravel(), np.float64)
    elif tn == "GaussianNB":
        m.kind, m.n_in = 2, est.theta_.shape[1]
        var = est.var_ if hasattr(est, "var_") else est.sigma_
        m.weights = arr(np.concatenate([est.theta_.ravel(), var.ravel()]), np.float64)
        m.biases = arr(np.log(est.class_prior_), np.float64)
    elif tn == "RandomForestClassifier":
        m.kind, m.n_in, m.n_trees = 3, est.n_features_in_, len(est.estimators_)
        off, L, R, F, T, P = [0], [], [], [], [], []
        for t in est.estimators_:
            tr = t.tree_
            L.append(tr.children_left); R.append(tr.children_right); F.append(np.maximum(tr.feature, 0)); T.append(tr.threshold)
            v = tr.value[:, 0, :]
            P.append(v[:, 1] / v.sum(axis=1))
            off.append(off[-1] + tr.node_count)
        m.tree_off = arr(off, np.int32)
        m.left, m.right, m.feature = arr(np.concatenate(L), np.int32), arr(np.concatenate(R), np.int32), arr(np.concatenate(F), np.int32)
        m.threshold, m.leaf_p1 = arr(np.concatenate(T), np.float64), arr(np.concatenate(P), np.float64)
    else:
        raise TypeError("unsupported estimator " + tn)
    return m, keep


def predict(est, X):
    m, keep = export_model(est)
    X = np.ascontiguousarray(X, dtype=np.float64)
    return np.array([lib().orc_predict(C.byref(m), X[i].ctypes.data_as(C.c_void_p)) for i in range(len(X))])


# ---- extract_features ---------------------------------------------------------------------------------

ERRORS = {-2: "output capacity", -3: "KeyError: read not in fastq", -4: "ValueError: numeric field", -5: "context / reference end",
          -6: "KeyError: model key", -7: "n diffs off"}


class OracleError(RuntimeError):
    pass


def fmt_float(x):
    """str(np.float64(x)) -- shortest round-trip repr."""
    return repr(float(x))


def fmt_prob(p):
    """str(np.round(p, 2)) (extract_contexts.py:207)."""
    return repr(float(lib().orc_np_round(float(p), 2)))


class Prepared(object):
    """Marshalled inputs of the oracle (marked reference, quality table, model) reusable across calls / threads."""
    pass


def prepare(fasta, quals, model=None, base="A", motif=None, positions=None):
    seqs = read_fasta(fasta) if isinstance(fasta, str) else fasta
    P = Prepared()
    P.names = list(seqs)
    P.carr = (_Contig * len(P.names))()
    P.keep = []
    for i, nm in enumerate(P.names):
        f, r = mark(seqs[nm], base, motif=motif, positions=positions, contig=nm)
        fb, rb, nb = f.encode(), r.encode(), nm.encode()
        P.keep += [fb, rb, nb]
        P.carr[i].name, P.carr[i].fwd, P.carr[i].rev, P.carr[i].len = nb, fb, rb, len(fb)
    qitems = sorted((kk.encode(), float(v)) for kk, v in quals.items())
    P.nq = len(qitems)
    P.qarr = (_Qual * max(len(qitems), 1))()
    for i, (kk, v) in enumerate(qitems):
        P.qarr[i].key, P.qarr[i].qual = kk, v
    P.keep.append(qitems)
    P.marr = (_Model * 2)()
    P.two = 0
    P.base = base
    P.have_model = model is not None
    if model is not None:
        if isinstance(model, dict):
            if base == "A":
                m0, k0 = export_model(model["MH"]); m1, k1 = export_model(model["MG"]); P.two = 1
                P.marr[0], P.marr[1] = m0, m1
                P.keep += k0 + k1
            else:
                m0, k0 = export_model(model["general"]); P.marr[0] = m0; P.keep += k0
        else:
            m0, k0 = export_model(model); P.marr[0] = m0; P.keep += k0
    return P


def extract(tsv_bytes, fasta, quals, k=6, skip_thresh=0, qual_thresh=0.0, model=None, base="A", motif=None,
            positions=None, cap=None, count_only=False, prepared=None):
    """Run the restated extract_features over a whole TSV (== reference with -t 1).

    Returns dict(rows=[str...], calls=[...], counters={...}) where rows are the `.diffs.<k>` lines.
    `model`: unpickled object (dict or bare estimator) or None (features only).  `prepared`: result of prepare()
    (then fasta / quals / model / base / motif / positions are ignored).
    """
    L = lib()
    P = prepared if prepared is not None else prepare(fasta, quals, model=model, base=base, motif=motif, positions=positions)
    names, carr, qarr, marr, two, base = P.names, P.carr, P.qarr, P.marr, P.two, P.base
    label_mod = "m6A" if base == "A" else "m" + base
    model = True if P.have_model else None
    qitems = range(P.nq)
    cap = cap or max(1024, len(tsv_bytes) // 64)
    calls = (_Call * cap)()
    n = L.orc_extract(C.c_char_p(tsv_bytes), C.c_int64(len(tsv_bytes)), carr, C.c_int32(len(names)), qarr, C.c_int64(len(qitems)),
                      C.c_int32(k), C.c_int32(skip_thresh), C.c_double(qual_thresh), marr, C.c_int32(two),
                      C.c_int32(1 if model is not None else 0), calls, C.c_int64(cap))
    if n < 0:
        raise OracleError(ERRORS.get(n, str(n)))
    if count_only:          # timing runs: the C restatement did all the work, skip the Python-side text rendering
        kinds = np.frombuffer(calls, dtype=np.uint8, count=n * C.sizeof(_Call)).reshape(n, C.sizeof(_Call))[:, 0] if n else np.zeros(0, np.uint8)
        return dict(rows=None, calls=None, counters=dict(observations=int((kinds == 0).sum())))
    rows, recs = [], []
    pos_set, multi, wskips, toomany = set(), set(), set(), set()
    for i in range(n):
        c = calls[i]
        read = tsv_bytes[c.read_off:c.read_off + c.read_len].decode()
        if c.kind == 1:
            toomany.add((read, c.mpos)); continue
        if c.kind == 2:
            multi.add((read, c.mpos)); continue
        feats = ["0" if (c.empty_mask >> j) & 1 else fmt_float(c.feat[j]) for j in range(k)] + [fmt_float(c.feat[k])]
        st = "-" if c.rev else "+"
        row = [names[c.chrom_cid], read, str(c.mpos), c.context.decode(), ",".join(feats), st]
        if model is not None:
            row += [label_mod if c.label else base, fmt_prob(c.prob)]
        rows.append("\t".join(row))
        recs.append(dict(chrom=names[c.chrom_cid], win_contig=names[c.win_cid], read=read, mpos=c.mpos, rev=bool(c.rev),
                         context=c.context.decode(), feat=[c.feat[j] for j in range(k + 1)], empty_mask=c.empty_mask,
                         prob=c.prob, label=c.label, model_sel=c.model_sel, n_empty=c.n_empty))
        pos_set.add(c.mpos)
        if c.n_empty > 0:
            wskips.add((read, c.mpos))
    return dict(rows=rows, calls=recs, counters=dict(observations=len(rows), positions=len(pos_set), multi=len(multi),
                                                     with_skips=len(wskips), too_many_skips=len(toomany)))


def aggregate(diffs_text, depth_thresh=15, mod_thresh=0.5, control=False):
    """make_bed.py aggregate_by_pos default mode -> list of BED row strings (first-seen order)."""
    L = lib()
    b = diffs_text.encode() if isinstance(diffs_text, str) else diffs_text
    cap = b.count(b"\n") + 2
    loci = (_Locus * cap)()
    n = L.orc_aggregate(C.c_char_p(b), C.c_int64(len(b)), loci, C.c_int64(cap))
    if n < 0:
        raise OracleError("malformed diffs row")
    out = []
    for i in range(n):
        lc = loci[i]
        key = [b[lc.key_off[j]:lc.key_off[j] + lc.key_len[j]].decode() for j in range(4)]
        frac = lc.meth / lc.depth          # np.mean of a 0/1 list: exact integer sum / n
        if lc.depth >= depth_thresh and ((not control and frac >= mod_thresh) or (control and frac < mod_thresh)):
            out.append("\t".join([key[0], key[1], str(int(key[1]) + 1), key[2], repr(float(frac)), key[3], str(lc.depth)]))
    return out


def aggregate_variants(diffs_text, depth_thresh=15, mod_thresh=0.5, control=False, pos_lines=None, verbose=False, gff=False):
    """make_bed.py aggregate_by_pos with -p / --vo / --gff (make_bed.py:67-164), restated row by row for small inputs.
    `pos_lines` is the text of the positions file (None without -p).  Uses scipy/numpy exactly where the reference does
    (stats.ttest_1samp :120, stats.sem :149, np.mean :143, np.round :127)."""
    import numpy as np
    from scipy import stats
    import warnings
    pos_set = None
    if pos_lines is not None:                                      # make_pos_set, :13-19
        pos_set = set()
        for line in pos_lines.splitlines(True):
            if len(line) > 3:
                pos_set.add(tuple(line.strip().split("\t")[:4]))
    pos_dict, values_dict, verbose_dict = {}, {}, {}
    for line in diffs_text.splitlines(True):
        f = line.split("\t")
        if len(f) == 8:
            csome, read, pos, context, values, strand, label, prob = f
        elif len(f) == 7:
            csome, read, pos, context, values, strand, label = f
            prob = ""
        else:
            raise OracleError("malformed diffs row")
        nextpos = str(int(pos) + 1)
        if (pos_set is not None and (csome, pos, nextpos, strand) not in pos_set) or context[int(len(context) / 2)] != "M":   # :84
            continue
        key = (csome, pos, nextpos, context, strand)
        if key not in pos_dict:
            pos_dict[key], values_dict[key], verbose_dict[key] = [], [], []
        if pos_set is not None:
            values_dict[key].append([float(v) for v in values.split(",")][:-1])      # :91
        pos_dict[key].append(1 if label[0] == "m" else 0)
        verbose_dict[key].append(prob.strip())
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if pos_set is not None:                                    # :115-127
            for key in values_dict:
                cols = list(zip(*values_dict[key]))
                pvals = []
                for col in cols:
                    t = stats.ttest_1samp(np.asarray(col, dtype=np.float64), 0)
                    pvals.append((t[1], t[0]))
                pval = (sum([-np.log10(x[0]) for x in pvals]), max([x[1] for x in pvals]))
                values_dict[key] = [np.round(x, 3) for x in [pval[1], pval[0]]]
        for key, lst in pos_dict.items():                          # :132-159
            frac = np.mean(lst)
            if pos_set is None:
                ok = len(lst) >= depth_thresh and ((not control and frac >= mod_thresh) or (control and frac < mod_thresh))
            else:
                ok = (key[0], key[1], key[2], key[4]) in pos_set
            if not ok:
                continue
            if gff:
                deets = "coverage=" + str(len(lst)) + ";context=" + key[3] + ";IPDRatio=5;frac=" + str(frac)
                if verbose:
                    probs = [float(x) for x in verbose_dict[key]]
                    se_95 = 2 * stats.sem(probs)
                    deets += ";fracLow=" + str(frac - se_95) + ";fracUp=" + str(frac + se_95) + ";identificationQv=" + str(int(100 * np.mean(probs)))
                out.append("\t".join([key[0], "kinModCall", "m6A", key[2], key[2], "10", key[4], ".", deets]))
            else:
                row = "\t".join(list(key)[:-1] + [str(frac)] + [key[-1]] + [str(len(lst))])
                if pos_set is not None:
                    row += "\t" + "\t".join([str(x) for x in values_dict[key]])
                if verbose:
                    row += "\t" + ",".join(verbose_dict[key])
                out.append(row)
    return out
