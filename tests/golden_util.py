import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
from make_golden import CASES, make_reads  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def check_extract(got, g):
    """got: flat dict from the oracle or the GPU; g: golden npz. Returns list of mismatching fields."""
    bad = []
    for f in ("hoco_l", "n_scm", "n_lrl", "n_n", "ho_l_rl", "n_nucl", "m_pos", "s_mer", "k_mer"):
        if f in got and not np.array_equal(np.asarray(got[f]), g[f]):
            bad.append(f)
    if not np.array_equal(digest(got["hoco_s"]), g["hoco_s_sha256"]):
        bad.append("hoco_s")
    if not np.array_equal(digest(got["ho_rl"]), g["ho_rl_sha256"]):
        bad.append("ho_rl")
    return bad


def check_scm(got, g):
    bad = []
    for f, gf in (("h", "scm_h"), ("s", "scm_s"), ("cov", "scm_cov")):
        if not np.array_equal(np.asarray(got[f]), g[gf]):
            bad.append(f)
    if not np.array_equal(digest(got["occ"]), g["scm_occ_sha256"]):
        bad.append("occ")
    if not np.array_equal(digest(got["k_mer_id"]), g["k_mer_id_sha256"]):
        bad.append("k_mer_id")
    return bad


def golden_arcs(g):
    """arcs of the reference graph with vertex ids mapped back to syncmer ids. The golden files hold
    them after asmg_finalize, whose asmg_arc_fix_symm (graph.c:205-235) flips the comp flag of an
    arc that is its own complement (v+ -> v-); make_syncmer_graph itself pushes it with comp = 0
    (syncasm.c:273-280), which is the a7 result compared here."""
    a = np.stack([g["arc_v"], g["arc_w"], g["arc_cov"], g["arc_comp"]], axis=1).astype(np.uint64)
    selfc = ((a[:, 1] ^ 1) == a[:, 0]) if len(a) else np.zeros(0, bool)
    a[selfc, 3] = 0
    order = np.lexsort((a[:, 2], a[:, 3], a[:, 1], a[:, 0]))
    return a[order]
