"""ctypes front-end to the CPU checkers. TEST INFRASTRUCTURE ONLY.

Loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package (oatk_b200).

  Oracle  -> oracle/liboracle.so   (our restatement, oracle/sync_oracle.c)
  Ref     -> oracle/_ref/libref.so (the unmodified reference, oracle/ref_shim.c)

Both return the same flat dictionary layout so tests can compare them with
each other and with the CUDA path field by field.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


def _p(a, t):
    return a.ctypes.data_as(t)


def build(force=False):
    """compile liboracle.so (always possible) and _ref/libref.so (only where /root/reference exists)"""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    src_t = max(os.path.getmtime(os.path.join(HERE, f)) for f in ("sync_oracle.c", "sync_oracle.h"))
    if not need and os.path.getmtime(os.path.join(HERE, "liboracle.so")) < src_t:
        need = True
    if need:
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(HERE, "_ref", "libref.so")
    if os.path.isdir("/root/reference") and (force or not os.path.exists(ref_so)
            or not os.path.exists(os.path.join(HERE, "_ref", "syncasm"))
            or os.path.getmtime(ref_so) < os.path.getmtime(os.path.join(HERE, "ref_shim.c"))):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def pack_reads(reads):
    """list of bytes -> (uint8 array of all bases, uint64 offsets n+1)"""
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return bases, off


def count_ambiguous(bases, off):
    ok = np.zeros(256, dtype=bool)
    ok[[0, 1, 2, 3]] = True
    for ch in b"ACGTUacgtu":
        ok[ch] = True
    bad = (~ok[bases[: int(off[-1])]]).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(bad)])
    return (cs[off[1:].astype(np.int64)] - cs[off[:-1].astype(np.int64)]).astype(np.uint32)


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.or_hash64.restype = C.c_uint64
        L.or_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.or_murmur64a.restype = C.c_uint64
        L.or_murmur64a.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.or_kmer_hash.restype = C.c_uint64
        L.or_kmer_hash.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        L.or_extract.restype = C.c_void_p
        L.or_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.or_db_free.argtypes = [C.c_void_p]
        L.or_totals.argtypes = [C.c_void_p, u64p]
        L.or_flatten.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        L.or_collect.restype = C.c_void_p
        L.or_collect.argtypes = [C.c_void_p, C.c_int]
        L.or_scm_free.argtypes = [C.c_void_p]
        L.or_stat.restype = C.c_int
        L.or_stat.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.or_analyze_count.restype = C.c_int
        L.or_analyze_count.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.or_debug_tie_suppressed.restype = C.c_uint64
        L.or_arcs.restype = C.c_uint64
        L.or_arcs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p]
        self.L = L

    def hash64(self, x, mask):
        return self.L.or_hash64(x, mask)

    def tie_suppressed(self):
        return int(self.L.or_debug_tie_suppressed())

    def murmur(self, b, seed=1234):
        return self.L.or_murmur64a(b, len(b), seed)

    def extract(self, bases, off, k, s):
        """returns (handle, flat dict). handle must be passed to free()."""
        n = len(off) - 1
        db = self.L.or_extract(bases.ctypes.data, off.ctypes.data, n, k, s)
        if not db:
            raise ValueError("bad k/s")
        return db, self._flat(db, n)

    def _flat(self, db, n):
        t = np.zeros(5, np.uint64)
        self.L.or_totals(db, _p(t, u64p))
        H, N, HB, NL, NN = (int(x) for x in t)
        f = dict(
            hoco_l=np.zeros(n, np.uint32), n_scm=np.zeros(n, np.uint32),
            n_lrl=np.zeros(n, np.uint32), n_n=np.zeros(n, np.uint32),
            hoco_s=np.zeros(HB + 1, np.uint8), ho_rl=np.zeros(H + 1, np.uint8),
            ho_l_rl=np.zeros(NL + 1, np.uint32), n_nucl=np.zeros(NN + 1, np.uint32),
            m_pos=np.zeros(N + 1, np.uint32), s_mer=np.zeros(N + 1, np.uint64), k_mer=np.zeros(N + 1, np.uint64))
        self.L.or_flatten(db, *[f[k].ctypes.data for k in
                                ("hoco_l", "n_scm", "n_lrl", "n_n", "hoco_s", "ho_rl", "ho_l_rl", "n_nucl",
                                 "m_pos", "s_mer", "k_mer")])
        for k_, ln in (("hoco_s", HB), ("ho_rl", H), ("ho_l_rl", NL), ("n_nucl", NN), ("m_pos", N), ("s_mer", N), ("k_mer", N)):
            f[k_] = f[k_][:ln]
        return f

    def collect(self, db, n_reads, hash_bits=64):
        """returns flat dict of the syncmer database, plus k_mer ids written back into db"""
        S = self.L.or_collect(db, hash_bits)
        if not S:
            return None

        class ScmT(C.Structure):
            _fields_ = [("n", C.c_uint64), ("n_occ", C.c_uint64), ("h", u64p), ("s", u64p), ("cov", u32p),
                        ("off", u64p), ("occ", u64p), ("smer_conflict", C.c_int)]
        st = ScmT.from_address(S)
        U, N = st.n, st.n_occ
        out = dict(
            h=np.ctypeslib.as_array(st.h, (U,)).copy(), s=np.ctypeslib.as_array(st.s, (U,)).copy(),
            cov=np.ctypeslib.as_array(st.cov, (U,)).copy(), off=np.ctypeslib.as_array(st.off, (U + 1,)).copy(),
            occ=np.ctypeslib.as_array(st.occ, (N,)).copy(), smer_conflict=int(st.smer_conflict), _handle=S)
        out["k_mer_id"] = self._flat(db, n_reads)["k_mer"]
        return out

    def stat(self, db):
        d = np.zeros(4, np.float64)
        i = np.zeros(8, np.int32)
        sc = np.zeros(1001, np.int64)
        kc = np.zeros(1001, np.int64)
        rc = self.L.or_stat(db, d.ctypes.data, i.ctypes.data, sc.ctypes.data, kc.ctypes.data)
        return rc, d, i, sc, kc

    def analyze_count(self, cnt, start_cnt=5):
        cnt = np.ascontiguousarray(cnt, dtype=np.int64)
        het = C.c_int(0)
        hom = self.L.or_analyze_count(len(cnt), start_cnt, cnt.ctypes.data, C.byref(het))
        return hom, het.value

    def arcs(self, db, scm, min_k_cov, a):
        n = self.L.or_arcs(db, scm["_handle"], min_k_cov, a, None)
        out = np.zeros((n, 4), np.uint64)
        if n:
            self.L.or_arcs(db, scm["_handle"], min_k_cov, a, out.ctypes.data)
        return out

    def free(self, db=None, scm=None):
        if scm is not None:
            self.L.or_scm_free(scm["_handle"])
        if db is not None:
            self.L.or_db_free(db)


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))


class Ref:
    """the unmodified reference (oracle/_ref/libref.so)"""

    def __init__(self):
        build()
        L = C.CDLL(os.path.join(HERE, "_ref", "libref.so"))
        L.ref_hash64.restype = C.c_uint64
        L.ref_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.ref_murmur64a.restype = C.c_uint64
        L.ref_murmur64a.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.ref_kmer_hash64.restype = C.c_uint64
        L.ref_kmer_hash64.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.ref_extract_mem.restype = C.c_void_p
        L.ref_extract_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.ref_extract_file.restype = C.c_void_p
        L.ref_extract_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_size_t]
        L.ref_sr_db_free.argtypes = [C.c_void_p]
        L.ref_scm_db_free.argtypes = [C.c_void_p]
        L.ref_scg_free.argtypes = [C.c_void_p]
        L.ref_n_reads.restype = C.c_uint64
        L.ref_n_reads.argtypes = [C.c_void_p]
        L.ref_totals.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_flatten.argtypes = [C.c_void_p] * 10
        L.ref_flatten_nnucl.argtypes = [C.c_void_p] * 3
        L.ref_stat.restype = C.c_int
        L.ref_stat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_collect.restype = C.c_void_p
        L.ref_collect.argtypes = [C.c_void_p]
        L.ref_scm_n.restype = C.c_uint64
        L.ref_scm_n.argtypes = [C.c_void_p]
        L.ref_scm_flatten.argtypes = [C.c_void_p] * 5
        L.ref_kmer_ids.argtypes = [C.c_void_p] * 2
        L.ref_make_graph.restype = C.c_void_p
        L.ref_make_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
        L.ref_unitig.argtypes = [C.c_void_p]
        L.ref_graph_n_vtx.restype = C.c_uint64
        L.ref_graph_n_vtx.argtypes = [C.c_void_p]
        L.ref_graph_n_arc.restype = C.c_uint64
        L.ref_graph_n_arc.argtypes = [C.c_void_p]
        L.ref_graph_arcs.argtypes = [C.c_void_p] * 2
        L.ref_graph_vtx_total.restype = C.c_uint64
        L.ref_graph_vtx_total.argtypes = [C.c_void_p]
        L.ref_graph_vtx.argtypes = [C.c_void_p] * 4
        L.ref_write_gfa.restype = C.c_int
        L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.ref_time_extract_count.restype = C.c_int
        L.ref_time_extract_count.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.L = L

    def hash64(self, x, mask):
        return self.L.ref_hash64(x, mask)

    def murmur(self, b, seed=1234):
        return self.L.ref_murmur64a(b, len(b), seed)

    def kmer_hash(self, hoco_s, start, k, rev):
        buf = np.ascontiguousarray(hoco_s)
        return self.L.ref_kmer_hash64(buf.ctypes.data, start << 1 | rev, k)

    def extract(self, bases, off, k, s):
        n = len(off) - 1
        db = self.L.ref_extract_mem(bases.ctypes.data, off.ctypes.data, n, k, s)
        return db, self._flat(db, n, count_ambiguous(bases, off))

    def extract_file(self, path, k, s, threads=1):
        db = self.L.ref_extract_file(path.encode(), k, s, threads, 0)
        return db, self._flat(db, int(self.L.ref_n_reads(db)), None)

    def _flat(self, db, n, n_n):
        t = np.zeros(4, np.uint64)
        self.L.ref_totals(db, t.ctypes.data)
        H, N, HB, NL = (int(x) for x in t)
        f = dict(
            hoco_l=np.zeros(n, np.uint32), n_scm=np.zeros(n, np.uint32), n_lrl=np.zeros(n, np.uint32),
            hoco_s=np.zeros(HB + 1, np.uint8), ho_rl=np.zeros(H + 1, np.uint8), ho_l_rl=np.zeros(NL + 1, np.uint32),
            m_pos=np.zeros(N + 1, np.uint32), s_mer=np.zeros(N + 1, np.uint64), k_mer=np.zeros(N + 1, np.uint64))
        self.L.ref_flatten(db, *[f[k].ctypes.data for k in
                                 ("hoco_l", "n_scm", "n_lrl", "hoco_s", "ho_rl", "ho_l_rl", "m_pos", "s_mer", "k_mer")])
        for k_, ln in (("hoco_s", HB), ("ho_rl", H), ("ho_l_rl", NL), ("m_pos", N), ("s_mer", N), ("k_mer", N)):
            f[k_] = f[k_][:ln]
        if n_n is not None:
            f["n_n"] = n_n
            nn = np.zeros(int(n_n.sum()) + 1, np.uint32)
            self.L.ref_flatten_nnucl(db, n_n.ctypes.data, nn.ctypes.data)
            f["n_nucl"] = nn[:-1]
        return f

    def analyze_count(self, cnt, start_cnt=5):
        cnt = np.ascontiguousarray(cnt, dtype=np.int64)
        het = C.c_int(0)
        self.L.ref_analyze_count.restype = C.c_int
        self.L.ref_analyze_count.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        hom = self.L.ref_analyze_count(len(cnt), start_cnt, cnt.ctypes.data, C.byref(het))
        return hom, het.value

    def stat(self, db, verbose=0):
        d = np.zeros(5, np.float64)
        i = np.zeros(8, np.int32)
        rc = self.L.ref_stat(db, verbose, d.ctypes.data, i.ctypes.data)
        return rc, d[1:], i

    def collect(self, db):
        S = self.L.ref_collect(db)
        if not S:
            return None
        U = int(self.L.ref_scm_n(S))
        t = np.zeros(4, np.uint64)
        self.L.ref_totals(db, t.ctypes.data)
        N = int(t[1])
        out = dict(h=np.zeros(U, np.uint64), s=np.zeros(U, np.uint64), cov=np.zeros(U, np.uint32),
                   occ=np.zeros(N, np.uint64), _handle=S)
        self.L.ref_scm_flatten(S, out["h"].ctypes.data, out["s"].ctypes.data, out["cov"].ctypes.data, out["occ"].ctypes.data)
        ids = np.zeros(N + 1, np.uint64)
        self.L.ref_kmer_ids(db, ids.ctypes.data)
        out["k_mer_id"] = ids[:N]
        out["off"] = np.concatenate([[0], np.cumsum(out["cov"].astype(np.uint64))]).astype(np.uint64)
        return out

    def graph(self, db, scm, min_k_cov, a):
        g = self.L.ref_make_graph(db, scm["_handle"], min_k_cov, a)
        return g

    def graph_dump(self, g):
        nv, na = int(self.L.ref_graph_n_vtx(g)), int(self.L.ref_graph_n_arc(g))
        arcs = np.zeros((na, 6), np.uint64)
        self.L.ref_graph_arcs(g, arcs.ctypes.data)
        tot = int(self.L.ref_graph_vtx_total(g))
        n = np.zeros(nv, np.uint64)
        fl = np.zeros(nv, np.uint64)
        ls = np.zeros(tot + 1, np.uint64)
        self.L.ref_graph_vtx(g, n.ctypes.data, fl.ctypes.data, ls.ctypes.data)
        return dict(arcs=arcs, vtx_n=n, vtx_flags=fl, vtx_lists=ls[:tot])

    def unitig(self, g):
        self.L.ref_unitig(g)

    def write_gfa(self, db, g, path):
        return self.L.ref_write_gfa(db, g, path.encode())

    def time_extract_count(self, path, k, s, threads):
        t = np.zeros(6, np.float64)
        rc = self.L.ref_time_extract_count(path.encode(), k, s, threads, t.ctypes.data)
        if rc:
            raise RuntimeError("reference run failed")
        return dict(sr_read_s=t[0], sr_db_stat_s=t[1], collect_s=t[2], syncmers=int(t[4]), distinct=int(t[5]))

    def free(self, db=None, scm=None, g=None):
        if g is not None:
            self.L.ref_scg_free(g)
        if scm is not None:
            self.L.ref_scm_db_free(scm["_handle"])
        if db is not None:
            self.L.ref_sr_db_free(db)
