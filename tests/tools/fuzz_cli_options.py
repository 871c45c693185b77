#!/usr/bin/env python
"""Random command lines for the `syncasm` binary of this repository against the reference's (oracle/_ref/syncasm): options,
shortened long options, `=` forms, clusters, stray values and file names that do not exist, with a `-h` dropped in so that
no run starts. Exit code, stdout and stderr must be equal. CPU only (everything compared ends before the device is
touched). Round 2: 1500 command lines, equal -- after it found that a file that cannot be opened ended our command with
three lines where the reference prints one.

  python tests/tools/fuzz_cli_options.py <seed> <number of command lines>"""
import os
import random, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OURS = os.path.join(ROOT, "oatk_b200", "host", "syncasm"); REF = os.path.join(ROOT, "oracle", "_ref", "syncasm")
opts = ["-k", "-s", "-c", "-a", "-D", "-t", "-o", "-v", "-V", "-h", "--max-bubble", "--max-tip", "--weak-cross", "--unzip-round",
        "--no-read-ec", "--version", "--help", "--verbose", "--threads", "--max", "--m", "--un", "--no", "--we", "--ve", "--he", "--v", "--t",
        "--max-b", "--max-t", "--unzip", "--weak", "--no-read", "--bogus", "-x", "-q", "--", "-", "-kx", "-k5", "-k5s3", "-hk", "-Vk", "-vv", "-t4k", "-o-"]
vals = ["5", "0", "-1", "abc", "3g", "12K", "1.5", "", "7M", "1e3", "0x10", "x.fa", "y.fa.gz", "99999999999999999999", "-5"]
def rnd_arg(r):
    a = r.choice(opts)
    if a.startswith("--") and len(a) > 2 and r.random() < 0.3:
        a += "=" + r.choice(vals)
    return a
def run(exe, args):
    p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, stdin=subprocess.DEVNULL, timeout=20)
    return p.returncode, p.stdout.decode().replace(exe, "syncasm"), p.stderr.decode().replace(exe, "syncasm")
r = random.Random(int(sys.argv[1])); n = int(sys.argv[2]); bad = 0
for it in range(n):
    args = []
    for _ in range(r.randint(1, 6)):
        args.append(rnd_arg(r))
        if r.random() < 0.6: args.append(r.choice(vals))
    # make sure no real run starts: a help request or an error must end the parse; -h goes to a random place
    args.insert(r.randint(0, len(args)), "-h")
    if "--" in args or "-" in args: continue          # "-" is standard input: a run would start
    try:
        a, b = run(OURS, args), run(REF, args)
    except subprocess.TimeoutExpired:
        print("timeout", args); continue
    started = "[M::syncasm]" in b[2] or "[M::sr_read" in b[2] or "failed to open" in b[2] or "[E::sstream" in b[2]
    if started: continue
    if a != b:
        bad += 1
        print("DIVERGENCE", args, "\n  ours:", a[0], repr(a[2][-200:]), "\n  ref :", b[0], repr(b[2][-200:]), flush=True)
print("done", n, "bad", bad)
