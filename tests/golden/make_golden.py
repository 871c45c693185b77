"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container, where /root/reference exists:  python tests/golden/make_golden.py
The inputs are regenerated from seeds by oatk_b200/synth.py; only the reference's outputs
are stored (hoco_s / ho_rl as SHA-256 digests, the small arrays verbatim).
"""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oatk_b200 import synth          # noqa: E402
from pyoracle import Ref, pack_reads  # noqa: E402

CASES = {
    # name: (generator, args, k, s, min_k_cov)
    "adv_k101_s11": ("adversarial", (3, 101, 11), 101, 11, 2),
    "adv_k64_s31": ("adversarial", (3, 64, 31), 64, 31, 2),
    "adv_k1001_s31": ("adversarial", (3, 1001, 31), 1001, 31, 2),
    "adv_k33_s31": ("adversarial", (5, 33, 31), 33, 31, 2),
    "hifi_k1001_s31": ("hifi", (42, 200000, 120, 15000, 0.001), 1001, 31, 3),
    "hifi_k501_s31": ("hifi", (7, 100000, 80, 12000, 0.002), 501, 31, 3),
}


def make_reads(gen, args):
    return synth.adversarial_reads(*args) if gen == "adversarial" else synth.hifi_reads(*args)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    R = Ref()
    for name, (gen, args, k, s, mkc) in CASES.items():
        reads = make_reads(gen, args)
        bases, off = pack_reads(reads)
        db, f = R.extract(bases, off, k, s)
        rc, d, i = R.stat(db)
        scm = R.collect(db)
        out = dict(hoco_l=f["hoco_l"], n_scm=f["n_scm"], n_lrl=f["n_lrl"], n_n=f["n_n"], ho_l_rl=f["ho_l_rl"], n_nucl=f["n_nucl"],
                   m_pos=f["m_pos"], s_mer=f["s_mer"], k_mer=f["k_mer"],
                   hoco_s_sha256=digest(f["hoco_s"]), ho_rl_sha256=digest(f["ho_rl"]),
                   stat_rc=np.array([rc]), stat_d=d, stat_i=i)
        if scm is not None:
            out.update(scm_h=scm["h"], scm_s=scm["s"], scm_cov=scm["cov"], scm_occ_sha256=digest(scm["occ"]),
                       k_mer_id_sha256=digest(scm["k_mer_id"]))
            g = R.graph(db, scm, mkc, 0.35)
            gd = R.graph_dump(g)
            # arcs with vertex ids mapped back to syncmer ids (asmg_cleanup renumbers the survivors)
            first = gd["vtx_lists"][np.concatenate([[0], np.cumsum(gd["vtx_n"])[:-1]]).astype(np.int64)] if len(gd["vtx_n"]) else np.zeros(0, np.uint64)
            arcs = gd["arcs"]
            v = (first[(arcs[:, 0] >> 1).astype(np.int64)] | (arcs[:, 0] & 1)) if len(arcs) else np.zeros(0, np.uint64)
            w = (first[(arcs[:, 1] >> 1).astype(np.int64)] | (arcs[:, 1] & 1)) if len(arcs) else np.zeros(0, np.uint64)
            out.update(arc_v=v, arc_w=w, arc_cov=arcs[:, 4] & 0x3FFFFFFF if len(arcs) else np.zeros(0, np.uint64),
                       arc_comp=(arcs[:, 4] >> 31) & 1 if len(arcs) else np.zeros(0, np.uint64), graph_n_vtx=np.array([len(gd["vtx_n"])]))
            R.free(g=g)
        R.free(db, scm)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "reads", len(reads), "syncmers", len(f["m_pos"]), "distinct", 0 if scm is None else len(scm["h"]))
    # unit vectors (SURVEY.md appendix A.4 and a few more), printed by the reference's own functions
    b = bytes((37 * i + 11) % 256 for i in range(256))
    mur = np.array([R.murmur(b[:n]) for n in (0, 1, 7, 8, 9, 11, 16, 250, 251, 252)], dtype=np.uint64)
    keys = [0, 1, 0x0123456789ABCDEF & ((1 << 62) - 1), (1 << 62) - 1, 12345678901234567 % (1 << 62)]
    h62 = np.array([R.hash64(x, (1 << 62) - 1) for x in keys], dtype=np.uint64)
    h22 = np.array([R.hash64(x, (1 << 22) - 1) for x in (0, 1, (1 << 22) - 1, 123456)], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "unit_vectors.npz"), murmur_len=np.array([0, 1, 7, 8, 9, 11, 16, 250, 251, 252]),
                        murmur=mur, hash62_key=np.array(keys, dtype=np.uint64), hash62=h62,
                        hash22_key=np.array([0, 1, (1 << 22) - 1, 123456], dtype=np.uint64), hash22=h22)


if __name__ == "__main__":
    main()
