"""field-by-field comparison of two flat result dictionaries with readable diagnostics"""
import numpy as np

EXTRACT_FIELDS = ("hoco_l", "n_scm", "n_lrl", "n_n", "hoco_s", "ho_rl", "ho_l_rl", "n_nucl", "m_pos", "s_mer", "k_mer")
SCM_FIELDS = ("h", "s", "cov", "off", "occ", "k_mer_id")


def diff(a, b, fields, tag=""):
    """returns a list of human-readable mismatch descriptions (empty = identical)"""
    out = []
    for f in fields:
        if f not in a or f not in b:
            continue
        x, y = np.asarray(a[f]), np.asarray(b[f])
        if x.shape != y.shape:
            out.append("%s%s: shape %s vs %s" % (tag, f, x.shape, y.shape))
            continue
        if not np.array_equal(x, y):
            bad = np.nonzero(x != y)[0]
            out.append("%s%s: %d of %d differ; first at %d: %s vs %s" % (
                tag, f, len(bad), x.size, bad[0], x[bad[:4]], y[bad[:4]]))
    return out


def per_read_report(got, exp, max_reads=5):
    """locate the reads where syncmer lists differ (for debugging)"""
    out = []
    go = np.concatenate([[0], np.cumsum(got["n_scm"].astype(np.int64))])
    eo = np.concatenate([[0], np.cumsum(exp["n_scm"].astype(np.int64))])
    n = min(len(got["n_scm"]), len(exp["n_scm"]))
    for r in range(n):
        g = got["m_pos"][go[r]:go[r + 1]]
        e = exp["m_pos"][eo[r]:eo[r + 1]]
        gs = got["s_mer"][go[r]:go[r + 1]]
        es = exp["s_mer"][eo[r]:eo[r + 1]]
        gk = got["k_mer"][go[r]:go[r + 1]]
        ek = exp["k_mer"][eo[r]:eo[r + 1]]
        if len(g) != len(e) or not (np.array_equal(g, e) and np.array_equal(gs, es) and np.array_equal(gk, ek)):
            sg_, se = set(g.tolist()), set(e.tolist())
            out.append("read %d hoco_l %d/%d: n %d vs %d; extra %s missing %s; smer_eq %s kmer_eq %s" % (
                r, got["hoco_l"][r], exp["hoco_l"][r], len(g), len(e), sorted(sg_ - se)[:6], sorted(se - sg_)[:6],
                len(g) == len(e) and np.array_equal(gs, es), len(g) == len(e) and np.array_equal(gk, ek)))
            if len(out) >= max_reads:
                break
    return out


# ---- the one input class on which the reference gives up ------------------------------------------------------
# (GC)n broken by two ambiguous bases in opposite phase: the k-mers behind them are the same k-mer on opposite strands, each
# is a syncmer through the same s-mer, but the stored s-mer code carries the other strand bit -- "identical kmers have
# different smers" (syncmer.c:1370-1376), four [E::process_kmer_cluster] lines and exit(EXIT_FAILURE). (Found by
# tests/tools/fuzz_oracle_vs_reference.py on a read whose homopolymer-compressed form is (GC)n.)
CONFLICT_K, CONFLICT_S = 129, 29
CONFLICT_READ = bytearray(b"GC" * 400)
CONFLICT_READ[200] = CONFLICT_READ[501] = ord("N")
CONFLICT_READ = bytes(CONFLICT_READ)

_REF_CONFLICT_SCRIPT = r"""
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from pyoracle import Ref, pack_reads
import parity
reads = [parity.CONFLICT_READ] if %d == 0 else eval(open(%r).read())
bases, off = pack_reads(reads)
ref = Ref()
db, _ = ref.extract(bases, off, parity.CONFLICT_K, parity.CONFLICT_S)
ref.collect(db)
print("survived")
"""


def reference_on_conflict(reads_file=None):
    """runs the unmodified reference (oracle/_ref/libref.so) on CONFLICT_READ (or the python list literal in reads_file) in
    a process of its own, because it exits; returns (exit code, [E::...] lines)"""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    script = _REF_CONFLICT_SCRIPT % (os.path.join(os.path.dirname(here), "oracle"), here, 1 if reads_file else 0, reads_file or "")
    p = subprocess.run([sys.executable, "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    return p.returncode, [l for l in p.stderr.decode().splitlines() if l.startswith("[E::")], p.stdout.decode()
