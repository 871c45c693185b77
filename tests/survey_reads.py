"""The read sets of SURVEY.md appendix A: generator A.1 (numpy PCG64 stream, seed / genome length / reads / read length / error rate)
and the answers the UNMODIFIED reference gave on them when the survey was written (A.3). They pin this repository to a run of the
reference that is independent of oracle/_ref: the FASTA bytes, the per-read syncmer dump and both GFA files of configs[0]."""
import hashlib
import numpy as np

READS10K = dict(seed=42, G=500000, N=10000, L=15000, err=0.001,
                fasta_md5="3a1b46df1d92eff064180920c7ee4ceb", fasta_bytes=150078462,
                hoco_total=112351586, syncmers=210576, distinct_kmers=123243,
                dump_md5="b051567ae7940860848cbf09c1407363",          # "R sid hoco_l n" + "m_pos s_mer k_mer" lines, k = 1001, s = 31
                first_read=(11296, 21, [(447, 1089221421968769627, 12466468101431059233), (997, 7955158109112731920, 17010454769244575422),
                                        (2937, 7955158109112731921, 13088768481959371214)]),
                utg_gfa_md5="cdc5f47d41f18e445af3b31fc4d5da23", final_gfa_md5="6bd1ab1264b4916009908230f3e54904",   # syncasm -k 1001 -s 31 -c 30 -t 8
                stat1=("number syncmers collected: 210576", "number uniqe smer: 7465; singletons: 5575", "average smer count: 28.208",
                       "smer peak_hom: 272; peak_het: 263", "number uniqe kmer: 123243; singletons: 120882", "average kmer count: 1.709",
                       "kmer peak_hom: 113; peak_het: 110", "average kmer space: -528.605"),
                after_ec=("number syncmers collected: 210358", "number uniqe kmer: 37977"),
                final=("number unitigs  : 17", "number syncmers : 664", "number arcs     : 0"))

READS80K = dict(seed=7, G=4000000, N=80000, L=15000, err=0.001,
                fasta_md5="36af7ad79e85c7fc6bbffae19a8928d0", fasta_bytes=1200708868,
                hoco_total=900590458, syncmers=1714278, distinct_kmers=1001918,
                dump_md5="756ca5945129d97276a62781a17a51cb",
                utg_gfa_md5="725ac164ab47eb550ecc0b16c18daf38", final_gfa_md5="72255ff76991504777e34831a7372b2e")   # syncasm -k 1001 -s 31 -c 30


def generate(seed, G, N, L, err):
    """list of reads (bytes) and the FASTA text, exactly as SURVEY.md A.1 writes them"""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    g2 = np.concatenate([genome, genome[:L]])                     # circular
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads, fa = [], []
    for i in range(N):
        p = int(rng.integers(0, G))
        r = g2[p:p + L].copy()
        if rng.integers(0, 2):
            r = (3 - r)[::-1]                                     # reverse complement
        ne = rng.binomial(L, err)
        if ne:
            pos = rng.integers(0, L, ne)
            kind = rng.integers(0, 3, ne)
            sub = pos[kind == 0]
            r[sub] = (r[sub] + rng.integers(1, 4, len(sub))) % 4
            dele = pos[kind == 1]
            keep = np.ones(L, bool)
            keep[dele] = False
            ins = np.sort(pos[kind == 2])
            r = r[keep]
            if len(ins):
                r = np.insert(r, np.minimum(ins, len(r)), rng.integers(0, 4, len(ins)).astype(np.uint8))
        b = A[r].tobytes()
        reads.append(b)
        fa.append(b">r%d\n" % i + b + b"\n")
    return reads, b"".join(fa)


def dump_md5(hoco_l, n_scm, m_pos, s_mer, k_mer):
    """md5 of the survey's dump harness output (A.2 ii): one R line per read, one line per syncmer, decimal"""
    h = hashlib.md5()
    so = np.concatenate([[0], np.cumsum(np.asarray(n_scm, dtype=np.int64))])
    for r in range(len(hoco_l)):
        h.update(b"R %d %d %d\n" % (r, int(hoco_l[r]), int(n_scm[r])))
        h.update(b"".join(b"%d %d %d\n" % (int(m_pos[j]), int(s_mer[j]), int(k_mer[j])) for j in range(so[r], so[r + 1])))
    return h.hexdigest()
