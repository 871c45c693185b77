#!/usr/bin/env python
"""Times sr_read_files (FASTA -> device pipeline -> the reference's per-read blocks) on a synthetic FASTA in /dev/shm,
once per pipeline slot count, each in a fresh process.  python tools/sr_read_bench.py [--reads 200000] [--slots 4,6,8]"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


def child(fa):
    from oatk_b200.host import build_host
    H = C.CDLL(build_host.build())
    H.sr_read_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_size_t]
    H.sr_db_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    H.sr_db_clean.argtypes = [C.c_void_p]
    files = (C.c_char_p * 1)(fa.encode())
    out = []
    for it in range(3):
        db = SrDb()
        H.sr_db_init(C.byref(db), 1001, 31)
        t0 = time.perf_counter()
        assert H.sr_read_files(C.byref(db), files, 1, 0) == 0
        out.append(time.perf_counter() - t0)
        H.sr_db_clean(C.byref(db))
    print("slots=%s sr_read_files: %s s" % (os.environ.get("OATK_SR_SLOTS", "default"), " ".join("%.3f" % t for t in out)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--slots", default="4,6,8")
    ap.add_argument("--child")
    args = ap.parse_args()
    if args.child:
        return child(args.child)
    import numpy as np
    rng = np.random.default_rng(3)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(tmp, "reads.fa")
    genome = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 10_000_000 + 15000)]
    with open(fa, "wb") as f:
        for i, p in enumerate(rng.integers(0, 10_000_000, args.reads)):
            f.write(b">r%d\n" % i)
            f.write(genome[p:p + 15000].tobytes())
            f.write(b"\n")
    for n in args.slots.split(","):
        env = dict(os.environ, OATK_SR_SLOTS=n)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", fa], env=env, check=True)
    os.unlink(fa)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
