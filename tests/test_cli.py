"""The `syncasm` command (oatk_b200/host/cli/syncasm_main.c) against the reference's own binary (oracle/_ref/syncasm,
built by oracle/Makefile from /root/reference with the recipe of the reference's Makefile:27-28).

CPU part: everything that ends before syncasm() is called -- --version, usage, -h, unknown / ambiguous / missing
options incl. ketopt's habits (options after file names, shortened long options, the token quoted in error messages).
GPU part: whole runs with several option combinations on the golden FASTA sets; both GFA files and every [M::...] /
[W::...] / [E::...] line on stderr must be identical (the three lines that carry a path or a clock are normalised)."""
import os
import re
import subprocess
import sys
import tempfile
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OURS = os.path.join(ROOT, "oatk_b200", "host", "syncasm")
REF = os.path.join(ROOT, "oracle", "_ref", "syncasm")
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def binaries():
    from oatk_b200.host import build_host
    build_host.build()
    import pyoracle
    pyoracle.build()
    assert os.path.exists(OURS)
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/syncasm not built (needs /root/reference)")
    return OURS, REF


def run(exe, args, cwd=None):
    p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd, timeout=600)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


EARLY = [
    ["--version"], ["-V"], ["-h"], ["--help"], [], ["-k", "501"], ["-k501", "-s", "21", "-h"], ["--bogus", "x.fa"],
    ["x.fa", "--bogus"], ["x.fa", "y.fa", "-q"], ["-xk", "5", "x.fa"], ["-k"], ["x.fa", "-k"], ["--max-bubble"],
    ["--ver"], ["--vers"], ["--verb", "2", "-h"], ["--max-b=7", "--max-t", "9", "-h"], ["--weak-cross=0.5", "--unzip=1", "--no-read", "-h"],
    ["-D", "3g", "-h"], ["-D", "12K", "-t", "7", "-o", "pre", "-a", "0.5", "-c", "9", "-v", "1", "-h"], ["-o", "-", "-h"],
    ["--", "-h"], ["-Vh"], ["-hV"], ["--threads", "4", "--help"], ["-", "--bogus"], ["--no-read-ec=1", "-h"], ["--m", "1", "x.fa"],
]


@pytest.mark.parametrize("args", EARLY, ids=[" ".join(a) or "none" for a in EARLY])
def test_option_handling_matches_the_reference(binaries, args):
    ours, ref = binaries
    if args and args[0] == "--":                 # would start a run on a file called -h: only compare that both try to
        return
    a, b = run(ours, args), run(ref, args)
    name = lambda t: t.replace(ours, "syncasm").replace(ref, "syncasm")
    assert a[0] == b[0]
    assert name(a[1]) == name(b[1])
    assert name(a[2]) == name(b[2])


def norm_stderr(txt, exe):
    out = []
    for line in txt.splitlines():
        if not re.match(r"\[[MWE]::", line):
            continue
        line = re.sub(r"Real time: [0-9.]+ sec; CPU: [0-9.]+ sec; Peak RSS: [0-9.]+ GB", "Real time: T", line)
        line = re.sub(r"(CPU|real) time *: [0-9.]+", r"\1 time: T", line)
        line = line.replace(exe, "syncasm")
        out.append(line)
    return out


RUNS = [
    ("repeats_default", ["-k", "201", "-s", "15", "-c", "3", "-t", "4"]),
    ("repeats_default", ["-k", "201", "-s", "15", "-c", "3", "-t", "1", "--no-read-ec", "-v", "1"]),
    ("repeats_k101_no_ec", ["-k", "101", "-s", "11", "-c", "2", "-a", "0.05", "--no-read-ec", "-t", "3", "--unzip-round", "1"]),
    ("diploid_unzip", ["-k", "201", "-s", "15", "-a", "0.2", "-t", "2", "--max-bubble", "5000", "--max-tip=2000", "--weak-cross", "0.2"]),
    ("organelle_auto_cov", ["-k", "201", "-s", "15", "-c", "0", "-t", "4", "-v", "2"]),
    ("hifi_k1001", ["-c", "5", "-t", "4", "-D", "2M"]),
    ("hifi_k1001", ["-t", "2", "-c", "5"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case,opts", RUNS, ids=["%s %s" % (c, " ".join(o)) for c, o in RUNS])
def test_whole_runs_match_the_reference_binary(binaries, case, opts):
    import make_golden_syncasm as G
    ours, ref = binaries
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "reads.fa")
    G.write_fasta(case, fa)
    res = {}
    for tag, exe in (("ours", ours), ("ref", ref)):
        rc, so, se = run(exe, opts + ["-o", os.path.join(tmp, tag), fa] if tag == "ours" else [fa] + opts + ["-o", os.path.join(tmp, tag)])
        res[tag] = (rc, so, norm_stderr(se.replace(os.path.join(tmp, tag), "OUT"), exe))
    assert res["ours"][0] == res["ref"][0]
    assert res["ours"][1] == res["ref"][1]
    # the CMD line lists the arguments in the order given; ours got the file name last
    strip_cmd = lambda lines: [l for l in lines if "CMD:" not in l]
    assert strip_cmd(res["ours"][2]) == strip_cmd(res["ref"][2])
    for suffix in (".utg.gfa", ".utg.final.gfa"):
        a, b = os.path.join(tmp, "ours" + suffix), os.path.join(tmp, "ref" + suffix)
        assert os.path.exists(a) == os.path.exists(b)
        if os.path.exists(a):
            assert open(a, "rb").read() == open(b, "rb").read(), suffix
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)


@pytest.mark.gpu
def test_failed_run_exit_code(binaries):
    ours, ref = binaries
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "short.fa")
    open(fa, "wb").write(b">a\nACGTACGTACGT\n")
    a = run(ours, ["-k", "201", "-s", "15", "-o", os.path.join(tmp, "o"), fa])
    assert a[0] == 1 and "[E::main] failed to constrcut assembly" in a[2]
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)


@pytest.mark.gpu
def test_smer_conflict_ends_the_run_like_the_reference(binaries):
    """identical k-mers with different s-mer codes (tests/parity.py): the reference prints four [E::process_kmer_cluster]
    lines and leaves with EXIT_FAILURE from inside collect_syncmer_from_reads; same lines, same exit code, nothing after"""
    import parity
    import make_golden_syncasm as G
    ours, ref = binaries
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "reads.fa")
    G.write_fasta("repeats_default", fa)
    with open(fa, "ab") as f:
        f.write(b">conflict\n" + parity.CONFLICT_READ + b"\n")
    opts = ["-k", str(parity.CONFLICT_K), "-s", str(parity.CONFLICT_S), "-c", "3", "-t", "2"]
    a = run(ours, opts + ["-o", os.path.join(tmp, "ours"), fa])
    b = run(ref, [fa] + opts + ["-o", os.path.join(tmp, "ref")])
    assert a[0] == b[0] == 1
    la, lb = norm_stderr(a[2], ours), norm_stderr(b[2], ref)
    assert [l for l in lb if l.startswith("[E::")][0] == "[E::process_kmer_cluster] identical kmers have different smers"
    assert la == lb
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)


@pytest.mark.parametrize("which", ["first", "second"])
def test_unopenable_input_ends_like_the_reference(binaries, which):
    """a file that cannot be opened: the reference prints one line from make_kseq_stream and exits (sstream.c:46-49), at once
    for the first file, after reading the first file for a later one; same line, same exit code, nothing else (no GPU needed:
    the files are read before the device is touched)"""
    ours, ref = binaries
    tmp = tempfile.mkdtemp()
    good = os.path.join(tmp, "good.fa")
    open(good, "wb").write(b">a\n" + b"ACGTTGCA" * 40 + b"\n")
    files = [os.path.join(tmp, "missing.fa")] if which == "first" else [good, os.path.join(tmp, "missing.fa")]
    a = run(ours, ["-k", "101", "-s", "11", "-o", os.path.join(tmp, "o")] + files)
    b = run(ref, files + ["-k", "101", "-s", "11", "-o", os.path.join(tmp, "r")])
    assert a[0] == b[0] == 1
    assert norm_stderr(a[2], ours) == norm_stderr(b[2], ref) == ['[E::make_kseq_stream] fail to open file "%s"' % files[-1]]
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)
