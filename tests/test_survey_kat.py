"""Known answers of the unmodified reference recorded in SURVEY.md appendix A.3 (an independent run, made when the survey was
written): the oracle must reproduce the reference's per-read syncmer dump of reads10k.fa bit for bit."""
import hashlib
import pyoracle
import survey_reads as S


def test_oracle_reproduces_the_reference_dump_of_reads10k():
    K = S.READS10K
    reads, fa = S.generate(K["seed"], K["G"], K["N"], K["L"], K["err"])
    assert len(fa) == K["fasta_bytes"] and hashlib.md5(fa).hexdigest() == K["fasta_md5"], "the generator no longer writes the survey's FASTA"
    bases, off = pyoracle.pack_reads(reads)
    O = pyoracle.Oracle()
    db, e = O.extract(bases, off, 1001, 31)
    assert int(e["hoco_l"].sum()) == K["hoco_total"] and len(e["m_pos"]) == K["syncmers"]
    hl, n, first = K["first_read"]
    assert int(e["hoco_l"][0]) == hl and int(e["n_scm"][0]) == n
    assert [(int(e["m_pos"][j]), int(e["s_mer"][j]), int(e["k_mer"][j])) for j in range(3)] == first
    assert S.dump_md5(e["hoco_l"], e["n_scm"], e["m_pos"], e["s_mer"], e["k_mer"]) == K["dump_md5"]
    scm = O.collect(db, len(reads))
    assert len(scm["h"]) == K["distinct_kmers"]
    O.free(db, scm)
