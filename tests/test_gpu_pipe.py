"""sg_pipe_run_host (chunked, multi-stream host path) must give exactly what one big batch gives."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
import parity

pytestmark = pytest.mark.gpu


def run_pipe(lib, bases, off, k, s, chunk_reads, n_slots):
    n = len(off) - 1
    total = int(off[-1])
    pipe = lib.Pipe(0, n_slots)
    capN = total // max(1, (k - s + 1) // 4) + 64 * n + 1024
    a = dict(
        hoco_l=np.zeros(n, np.uint32), n_scm=np.zeros(n, np.uint32),
        hoco_s_off=np.zeros(n + 1, np.uint64), ho_rl_off=np.zeros(n + 1, np.uint64), scm_off=np.zeros(n + 1, np.uint64),
        hoco_s_buf=np.zeros(total // 4 + 16 * n + 64, np.uint8), ho_rl_buf=np.zeros(total + 16 * n + 64, np.uint8),
        m_pos=np.zeros(capN, np.uint32), s_mer=np.zeros(capN, np.uint64), k_mer=np.zeros(capN, np.uint64),
        amb_sid=np.zeros(total + 1, np.uint32), amb_pos=np.zeros(total + 1, np.uint32),
        lrl_sid=np.zeros(total // 256 + 8, np.uint32), lrl_idx=np.zeros(total // 256 + 8, np.uint32), lrl_val=np.zeros(total // 256 + 8, np.uint32))
    o = lib.ExtractOut()
    for name, _ in lib.ExtractOut._fields_:
        setattr(o, name, a[name].ctypes.data)
    caps = lib.PipeCaps(capN, a["hoco_s_buf"].size, a["ho_rl_buf"].size, total + 1, total // 256 + 8)
    z = pipe.run_host(bases.ctypes.data, off.ctypes.data, n, k, s, chunk_reads, o, caps)
    N = z.n_syncmers
    hl = a["hoco_l"].astype(np.int64)
    f = dict(hoco_l=a["hoco_l"], n_scm=a["n_scm"], m_pos=a["m_pos"][:N], s_mer=a["s_mer"][:N], k_mer=a["k_mer"][:N],
             n_lrl=np.bincount(a["lrl_sid"][:z.n_long_runs], minlength=n).astype(np.uint32),
             n_n=np.bincount(a["amb_sid"][:z.n_ambiguous], minlength=n).astype(np.uint32),
             ho_l_rl=a["lrl_val"][:z.n_long_runs], n_nucl=a["amb_pos"][:z.n_ambiguous],
             hoco_s=lib._unpad(a["hoco_s_buf"], a["hoco_s_off"], (hl + 3) // 4), ho_rl=lib._unpad(a["ho_rl_buf"], a["ho_rl_off"], hl))
    assert int(a["scm_off"][-1]) == N
    return pipe, f


@pytest.mark.parametrize("chunk,slots", [(7, 3), (64, 2), (1000, 3), (1, 1)])
def test_pipe_equals_single_batch(gpu_ctx, oracle, chunk, slots):
    from oatk_b200 import lib
    k, s = 501, 31
    reads = synth.hifi_reads(5, 120000, 150, 11000, 0.001) + synth.adversarial_reads(3, k, s)
    if chunk == 1:
        reads = reads[:40]
    bases, off = pack_reads(reads)
    db, exp = oracle.extract(bases, off, k, s)
    pipe, got = run_pipe(lib, bases, off, k, s, chunk, slots)
    d = parity.diff(got, exp, parity.EXTRACT_FIELDS)
    assert not d, "\n".join(d + parity.per_read_report(got, exp))
    # statistics, database and arcs from the master batch
    st = pipe.master.stat()
    rc, dd, ii, sc, kc = oracle.stat(db)
    assert np.array_equal(np.array(st.smer_cnts[:], np.int64), sc) and np.array_equal(np.array(st.kmer_cnts[:], np.int64), kc)
    assert st.gap_sum / st.n_gaps == dd[1]
    pipe.master.count()
    scm = pipe.master.count_download()
    es = oracle.collect(db, len(reads))
    assert not parity.diff(scm, es, parity.SCM_FIELDS)
    arcs = pipe.master.arcs(2, 0.35)
    assert np.array_equal(arcs, oracle.arcs(db, es, 2, 0.35))
    oracle.free(db, es)
    pipe.close()
