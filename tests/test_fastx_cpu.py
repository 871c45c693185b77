"""f4 on the CPU: the native FASTA/FASTQ reader of the host layer (oatk_b200/host/fastx_gpu.c) against the
reference's own reader (sstream_read -> kseq_read over zlib, driven like sr_read drives it) on files that
exercise the record grammar: multi-line records, FASTQ, CRLF, empty lines, header characters inside quality
strings, truncated records, several files, gzip, the -D cap. Needs oracle/_ref/libref.so; no GPU."""
import ctypes as C
import gzip
import os
import tempfile
import numpy as np
import pytest


class FastX(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("off", C.POINTER(C.c_uint64)), ("names", C.POINTER(C.c_char_p)),
                ("n", C.c_uint64), ("m", C.c_uint64), ("n_bases", C.c_uint64), ("m_bases", C.c_uint64), ("limit_reached", C.c_int)]


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.fastx_load.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_uint64, C.POINTER(FastX)]
    L.fastx_free.argtypes = [C.POINTER(FastX)]
    return L


def ours(host, paths, max_bases=0):
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    x = FastX()
    assert host.fastx_load(arr, len(paths), max_bases, C.byref(x)) == 0
    n = int(x.n)
    off = [int(x.off[i]) for i in range(n + 1)]
    raw = C.string_at(x.bases, off[n]) if n else b""
    out = [(bytes(x.names[i]), raw[off[i]:off[i + 1]]) for i in range(n)]
    lim = bool(x.limit_reached)
    host.fastx_free(C.byref(x))
    return out, lim


def theirs(ref, paths, max_bases=0):
    L = ref.L
    L.ref_parse_files.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64)),
                                  C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    b, o, nm, no, n = C.c_void_p(), C.POINTER(C.c_uint64)(), C.c_void_p(), C.POINTER(C.c_uint64)(), C.c_uint64()
    assert L.ref_parse_files(arr, len(paths), max_bases, C.byref(b), C.byref(o), C.byref(nm), C.byref(no), C.byref(n)) == 0
    n = int(n.value)
    off = [int(o[i]) for i in range(n + 1)]
    noff = [int(no[i]) for i in range(n + 1)]
    raw = C.string_at(b, off[n]) if n else b""
    names = C.string_at(nm, noff[n]) if n else b""
    out = [(names[noff[i]:noff[i + 1]], raw[off[i]:off[i + 1]]) for i in range(n)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for p in (b, C.cast(o, C.c_void_p), nm, C.cast(no, C.c_void_p)):
        libc.free(p)
    return out


def _write(d, name, data, gz=False):
    p = os.path.join(d, name)
    if gz:
        with gzip.open(p, "wb") as f:
            f.write(data)
    else:
        with open(p, "wb") as f:
            f.write(data)
    return p


CASES = {
    "fasta_one_line": b">r1 comment here\nACGTACGT\n>r2\nGGGG\n",
    "fasta_multi_line": b">r1\nACGT\nACGT\n\nTT\n>r2\tx\nA\nC\n",
    "fasta_no_final_newline": b">r1\nACGT\n>r2\nGG",
    "fasta_crlf": b">r1 c\r\nACGT\r\nAC\r\n>r2\r\nG\r\n",
    "fasta_leading_garbage": b"junk line\nmore>r1\nACGT\n",
    "fasta_empty_records": b">r1\n>r2\n\n>r3\nAC\n>\nGG\n",
    "fasta_lower_iupac": b">r1\nacgtnNRYKM\nacgu\n",
    "fastq_plain": b"@q1 c\nACGT\n+\nIIII\n@q2\nGG\n+q2\n##\n",
    "fastq_multi_line": b"@q1\nACGT\nACGT\n+\nIIII\nIIII\n@q2\nA\n+\nI\n",
    "fastq_at_in_quality": b"@q1\nACGT\n+\n@III\n@q2\nACGT\n+\n>>>>\n@q3\nAC\n+\nII\n",
    "fastq_crlf": b"@q1\r\nACGT\r\n+\r\nIIII\r\n@q2\r\nAC\r\n+\r\nII\r\n",
    "fastq_truncated_quality": b"@q1\nACGT\n+\nIIII\n@q2\nACGTAC\n+\nIII\n",
    "fastq_missing_quality": b"@q1\nACGT\n+\nIIII\n@q2\nACGT\n+",
    "fastq_long_quality": b"@q1\nACGT\n+\nIIIIII\n@q2\nAC\n+\nII\n",
    "fastq_empty_sequence": b"@q1\n+\n\n@q2\nAC\n+\nII\n",
    "mixed": b">r1\nACGT\n@q1\nGGGG\n+\nIIII\n>r2\nTT\n",
    "plus_line_in_fasta": b">r1\nAC\n+GT\nACGT\n>r2\nTT\n",
    "header_only_at_end": b">r1\nACGT\n>",
    "name_at_end": b">r1\nACGT\n>r2",
    "empty": b"",
    "only_newlines": b"\n\n\n",
    "lone_cr_line": b">r1\nAC\n\r\nGT\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_grammar_matches_kseq(host, ref, name):
    with tempfile.TemporaryDirectory() as d:
        p = _write(d, name + ".fx", CASES[name])
        got, _ = ours(host, [p])
        assert got == theirs(ref, [p])
        pz = _write(d, name + ".fx.gz", CASES[name], gz=True)
        got, _ = ours(host, [pz])
        assert got == theirs(ref, [pz])


def test_buffer_boundaries(host, ref):
    """kseq learns about the end of file from a short read of its 16 KB buffer: sizes around multiples of it"""
    rng = np.random.default_rng(1)
    with tempfile.TemporaryDirectory() as d:
        for target in (16384, 32768, 16383, 16385, 49152):
            for tail in (b"", b"\n", b"\r", b"\r\n", b">", b"\n>", b"\n+", b"\n+\n"):
                body = b">r1\n"
                seq = bytes(b"ACGT"[i] for i in rng.integers(0, 4, target - len(body) - len(tail)))
                data = body + seq + tail
                assert len(data) == target
                p = _write(d, "b.fa", data)
                got, _ = ours(host, [p])
                assert got == theirs(ref, [p]), (target, tail)


def test_random_files_and_caps(host, ref):
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for fi in range(3):
            recs = []
            for i in range(int(rng.integers(1, 60))):
                L = int(rng.integers(0, 400))
                seq = bytes(b"ACGTNacgtn"[j] for j in rng.integers(0, 10, L))
                w = int(rng.integers(1, 90))
                lines = [seq[a:a + w] for a in range(0, L, w)] or [b""]
                eol = b"\r\n" if rng.integers(0, 4) == 0 else b"\n"
                if rng.integers(0, 2):
                    recs.append(b">f%d_%d some comment" % (fi, i) + eol + eol.join(lines) + eol)
                else:
                    q = bytes(rng.integers(33, 75, L, dtype=np.uint8).tolist())
                    qlines = [q[a:a + w] for a in range(0, L, w)] or [b""]
                    recs.append(b"@f%d_%d" % (fi, i) + eol + eol.join(lines) + eol + b"+" + eol + eol.join(qlines) + eol)
            paths.append(_write(d, "f%d.fx%s" % (fi, ".gz" if fi == 1 else ""), b"".join(recs), gz=(fi == 1)))
        want = theirs(ref, paths)
        got, lim = ours(host, paths)
        assert got == want and not lim and len(got) > 3
        total = sum(len(s) for _, s in want)
        for cap in (1, total // 3, total - 1, total, total + 1):
            got, lim = ours(host, paths, cap)
            assert got == theirs(ref, paths, cap), cap
            assert lim == (cap <= total)


def _big_fasta(rng, n_rec, poison=None):
    recs = []
    for i in range(n_rec):
        L = int(rng.integers(0, 30000))
        seq = np.frombuffer(b"ACGTacgtN>", dtype=np.uint8)[rng.integers(0, 10, L)].tobytes()   # '>' inside lines is sequence data
        w = int(rng.integers(60, 20000))
        eol = b"\r\n" if i % 7 == 0 else b"\n"
        lines = [seq[a:a + w] for a in range(0, L, w)] or [b""]
        lines = [ln if not ln.startswith(b">") else b"A" + ln for ln in lines]                   # ... but not at a line start
        recs.append(b">rec%d extra words" % i + eol + eol.join(lines) + eol + (b"\n" if i % 5 == 0 else b""))
        if poison is not None and i == n_rec // 2:
            recs.append(poison)
    return b"".join(recs)


@pytest.mark.parametrize("poison", [None, b"+odd line\n", b"@q\nACGT\n+\nIIII\n"])
def test_parallel_fasta_path(host, ref, poison, monkeypatch):
    monkeypatch.setenv("OATK_FASTX_THREADS", "16")      # 32 pieces whatever the machine
    _parallel_case(host, ref, poison)


def _parallel_case(host, ref, poison):
    """files above 8 MB that are plain FASTA are cut at header lines and parsed by several threads; a line that starts
    with '+' or '@' anywhere sends the file back to the sequential parser. Either way: the reference's records."""
    rng = np.random.default_rng(3)
    with tempfile.TemporaryDirectory() as d:
        data = _big_fasta(rng, 900, poison)
        assert len(data) > (8 << 20)
        p = _write(d, "big.fa", data)
        want = theirs(ref, [p])
        got, lim = ours(host, [p])
        assert got == want and not lim
        total = sum(len(s) for _, s in want)
        for cap in (total // 2, total):
            got, lim = ours(host, [p], cap)
            assert got == theirs(ref, [p], cap) and lim
        # two big files in a row, the second gzipped
        p2 = _write(d, "big2.fa.gz", data[:len(data) // 2], gz=True)
        got, _ = ours(host, [p, p2])
        assert got == theirs(ref, [p, p2])


def test_fuzz_against_kseq(host, ref):
    """byte soups made of the characters the grammar cares about: whatever kseq makes of them, so must we"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    alphabet = [b">", b"@", b"+", b"A", b"C", b"g", b"N", b"\r", b" ", b"\t", b"I", b"#"]
    line = st.lists(st.sampled_from(alphabet), min_size=0, max_size=12).map(b"".join)
    doc = st.tuples(st.lists(line, min_size=0, max_size=14), st.booleans()).map(lambda t: b"\n".join(t[0]) + (b"\n" if t[1] else b""))
    d = tempfile.mkdtemp()
    p = os.path.join(d, "fuzz.fx")

    @settings(max_examples=600, deadline=None, suppress_health_check=list(HealthCheck))
    @given(doc)
    def run(data):
        with open(p, "wb") as f:
            f.write(data)
        got, _ = ours(host, [p])
        assert got == theirs(ref, [p]), data

    run()
    os.unlink(p)
    os.rmdir(d)
