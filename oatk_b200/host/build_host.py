"""Builds liboatk_gpu.so: the C host layer that mirrors the reference's syncmer.h API over libsyncgpu."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    if os.environ.get("OATK_HOST_LIB"):          # a prebuilt variant (e.g. an AddressSanitizer build for the test-suite)
        return os.environ["OATK_HOST_LIB"]
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".c")]
    if not srcs:
        return None
    out = os.path.join(HERE, "liboatk_gpu.so")
    deps = srcs + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        build_cli()
        return out
    lib = os.path.join(HERE, "..")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-o", out] + srcs +
                          ["-I" + os.path.join(HERE, "..", "..", "include"), "-L" + lib, "-lsyncgpu",
                           "-Wl,-rpath,$ORIGIN/..", "-lm", "-lpthread", "-lz"])
    build_cli(force=True)
    return out


def build_cli(force=False):
    """the `syncasm` command (reference run_syncasm.c main): cli/syncasm_main.c over liboatk_gpu.so"""
    src = os.path.join(HERE, "cli", "syncasm_main.c")
    exe = os.path.join(HERE, "syncasm")
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(src):
        return exe
    subprocess.check_call(["gcc", "-O2", "-Wall", "-o", exe, src, "-I" + os.path.join(HERE, "..", "..", "include"),
                           "-L" + HERE, "-loatk_gpu", "-L" + os.path.join(HERE, ".."), "-lsyncgpu",
                           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/..", "-lm", "-lpthread", "-lz"])
    return exe
