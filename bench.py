#!/usr/bin/env python
"""bench.py -- HiFi bases/second through syncmer extract + count.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path (sg_extract + sg_stat + sg_count = reference
sr_read_analysis_thread + sr_db_stat + collect_syncmer_from_reads) over one batch of
synthetic reads that is already resident in HBM: BASELINE.json configs[1], 1 M x 15 kb
reads per GPU at k=1001 s=31. `value` is whole-job raw bases per second; `e2e` is the same
pass through the C ABI with HOST buffers (pinned), host<->device copies inside the timed
region. `--impl reference` times the unmodified reference (oracle/_ref/libref.so) on the
host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, S = 1001, 31
READ_LEN = 15000
ERR = 1e-3


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def write_fasta(path, bases_u8, n_reads, read_len):
    import numpy as np
    hdr = np.frombuffer(b"".join(b">r%07d\n" % i for i in range(n_reads)), dtype=np.uint8).reshape(n_reads, 10)
    rec = np.empty((n_reads, 10 + read_len + 1), dtype=np.uint8)
    rec[:, :10] = hdr
    rec[:, 10:10 + read_len] = bases_u8.reshape(n_reads, read_len)
    rec[:, -1] = 10
    rec.tofile(path)


def host_sample_reads(n_reads, seed):
    """the bench workload's generator on the CPU path: numpy, same error model (oatk_b200/synth.py)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    G = 50_000_000
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = np.empty((n_reads, READ_LEN), dtype=np.uint8)
    start = rng.integers(0, G - READ_LEN - 64, n_reads)
    for i in range(n_reads):
        r = genome[start[i]:start[i] + READ_LEN].copy()
        ne = rng.binomial(READ_LEN, ERR)
        pos = rng.integers(0, READ_LEN, ne)
        r[pos] = (r[pos] + rng.integers(1, 4, ne)) % 4       # substitutions only: keeps the length fixed
        if rng.integers(0, 2):
            r = (3 - r)[::-1]
        out[i] = acgt[r]
    return out.reshape(-1)


def cpu_reference_run(bases_u8, n_reads, threads):
    """times the unmodified reference (sr_read + sr_db_stat + collect) on a FASTA of the sample"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if not pyoracle.have_ref():
        pyoracle.build()
    if not pyoracle.have_ref():
        return None
    R = pyoracle.Ref()
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "bench_sample_%d.fa" % os.getpid())
    write_fasta(path, bases_u8, n_reads, READ_LEN)
    try:
        t = R.time_extract_count(path, K, S, threads)
    finally:
        os.unlink(path)
    tot = t["sr_read_s"] + t["sr_db_stat_s"] + t["collect_s"]
    return {"seconds": tot, "stages": t, "bases": n_reads * READ_LEN}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_reads = min(10000 * threads, 160000)
    bases = host_sample_reads(n_reads, 1)
    times = []
    last = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(bases, n_reads, threads)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so missing and /root/reference absent"}))
            return 0
        if i >= args.warmup:
            times.append(r["seconds"])
        last = r
    sec = sum(times) / len(times)
    val = n_reads * READ_LEN / sec
    sample = "%d x %d b reads (%.2f Gbases) per step, reference sr_read(kseq parse + extract, -t %d) + sr_db_stat + collect" % (
        n_reads, READ_LEN, n_reads * READ_LEN / 1e9, threads)
    line = {
        "impl": "reference", "metric": "HiFi bases/sec syncmer-extract+count", "value": val, "unit": "bases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "bases/s", "cores": threads, "kind": "reference", "sample": sample,
                         "stages_s": last["stages"]},
        "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, n_reads=None, workload="uniform", n_planted=0):
    extra = "" if workload == "uniform" else "; %d reads per GPU carry a 2-15 kb tandem array (period 2/3/6=TTAGGG/37/171)" % n_planted
    return {"workload": "syncmer extract+count, %s x %d b synthetic HiFi reads per GPU, k=%d s=%d (BASELINE.json configs[1])%s" % (
                "1M" if n_reads in (None, 1000000) else str(n_reads), READ_LEN, K, S, extra),
            "input_class": workload,
            "k": K, "s": S, "read_len": READ_LEN, "error_rate": ERR, "genome_len": 50_000_000,
            "reads_per_gpu": n_reads or 1000000, "parallelism": "reads sharded by record, %d GPU(s)" % n_gpus,
            "l2": "inputs (15 GB per step) larger than L2; no flush needed"}


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU (default: the configs[1] workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-full-e2e", action="store_true", help="skip the second e2e measurement (run lengths downloaded as well)")
    ap.add_argument("--k", type=int, default=K, help="k-mer size (configs[4] sweep: 501, 1001, 2001)")
    ap.add_argument("--err", type=float, default=ERR, help="per-base error rate of the synthetic reads")
    ap.add_argument("--workload", default="uniform", choices=("uniform", "repeats"),
                    help="uniform: configs[1] as is; repeats: the same with every 100th read carrying a 2-15 kb tandem array "
                         "(period 2, 3, TTAGGG, 37, 171)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] lines (k = 501 and 2001 on the same reads)")
    ap.add_argument("--no-config3", action="store_true", help="skip the configs[3] line (10 M reads over 8 GPUs; runs only at 8 GPUs)")
    ap.add_argument("--no-whole", action="store_true", help="skip the configs[2] line (whole syncasm command on 200 k reads; one GPU only)")
    ap.add_argument("--whole-reads", type=int, default=200000)
    ap.add_argument("--no-numa-bind", action="store_true", help="leave the process where the launcher put it (default: the CPUs and the "
                    "memory node next to this rank's GPU at N > 1, the memory node only at N = 1, where the CPU baseline needs every core)")
    ap.add_argument("--config3-reads", type=int, default=0, help="run the configs[3] line with this many reads in all at any N > 1 (default: 10 M, at 8 GPUs only)")
    args = ap.parse_args()
    globals()["K"], globals()["ERR"] = args.k, args.err
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from oatk_b200 import lib, synth_gpu

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsyncgpu has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if not args.no_numa_bind:
        # one process per GPU: pinned host buffers on the GPU's own NUMA node (sg_host_bind_near_device), before any is allocated
        node, ncpu = lib.bind_host_near_device(local, cpus=world > 1, memory=True)
        numa = {"node": node, "cpus_bound": ncpu}
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    n_reads = args.reads

    bases, off = synth_gpu.hifi_reads_gpu(1000 + rank, 50_000_000, n_reads, READ_LEN, ERR, dev, genome_seed=1)
    total = n_reads * READ_LEN
    n_planted = 0
    if args.workload == "repeats":
        n_planted = synth_gpu.plant_repeats_gpu(bases, n_reads, READ_LEN, 100, 77 + rank, dev)
    torch.cuda.synchronize()

    ctx = lib.Context(local)          # launches on the legacy default stream = torch's current stream
    batch = lib.Batch(ctx)
    batch.set_sid_base(rank * n_reads)
    comm = None
    mgpu = None
    if world > 1:
        # parity first: a seeded read set (HiFi-like + adversarial + tandem arrays) through the same C-side exchange,
        # checked on rank 0 against the CPU oracle run on the WHOLE set; the result rides on the JSON line
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mgpu_parity
        try:
            mgpu = mgpu_parity.run(dist, rank, world, local)
        except Exception as e:                       # a failed check must not take the timing run with it: it is reported
            mgpu = {"ok": False, "world": world, "error": repr(e)[:300]}
        comm = lib.Comm(ctx, world, rank, share_unique_id(torch, dist, lib, rank, dev))

    def step(k=None, rd=None):
        b_, o_, n_, t_ = rd if rd is not None else (bases, off, n_reads, total)
        batch.set_reads_device(b_.data_ptr(), o_.data_ptr(), n_, t_)
        batch.extract(K if k is None else k, S)
        if comm is not None:
            comm.exchange_tuples(batch)      # sg_tuples_partition -> NCCL all-to-all-v -> sg_tuples_adopt, in C
        st = batch.stat()
        batch.count()
        if comm is not None:
            comm.return_ids(batch)           # global ids back to the GPU that holds the read
        return st

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    sizes = batch.extract_sizes()
    csz = batch.count_sizes()

    # ---- device-resident timing ----
    ctx.enable_timing(True)
    ctx.timings()
    stage_ms = {}
    stage_launch = {}
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launches()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
        for name, (ms, ln) in ctx.timings().items():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms
            stage_launch[name] = stage_launch.get(name, 0) + ln
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches = ctx.launches() - l0
    ex_bytes = None if comm is None else comm.bytes_sent() // max(1, W + args.steps)
    clocks = sampler.stop() if sampler else None
    ms_step = e0.elapsed_time(e1) / args.steps
    by_rank = None
    if dist is not None:
        t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank's own stage times (the JSON line carries rank 0's): the spread shows where ranks wait for each other
        names = sorted(stage_ms)
        mine = torch.tensor([stage_ms[n_] / args.steps for n_ in names] + [ms_step], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        by_rank = {n_: [round(float(r_[i]), 3) for r_ in allr] for i, n_ in enumerate(names) if float(max(r_[i] for r_ in allr)) > 0.05}
        by_rank["step"] = [round(float(r_[-1]), 3) for r_ in allr]
        ms_step = float(t.item())
    ctx.enable_timing(False)
    value = world * total / (ms_step * 1e-3)

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, lib, ctx, batch, bases, off, n_reads, total, sizes, csz, max(1, min(args.steps, 3)), dist, dev, world, rank)
        if world == 1 and not args.no_full_e2e:
            ef = run_e2e(torch, lib, ctx, batch, bases, off, n_reads, total, sizes, csz, 2, dist, dev, world, rank, full=True)
            e2e["full_download"] = {k_: ef[k_] for k_ in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "steps", "run_lengths")}
    ids_check = None
    if world > 1:
        ids_check = check_global_ids(torch, dist, batch, comm, rank, world, dev, n_reads)

    def timed(k, rd, steps):
        """device-resident throughput of one more configuration, same protocol as the headline (barrier, CUDA events, max over ranks)"""
        n_, t_ = (rd[2], rd[3]) if rd is not None else (n_reads, total)
        for _ in range(3):
            step(k, rd)
        ctx.enable_timing(True)
        ctx.timings()
        sm = {}
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step(k, rd)
            for name, (ms, _ln) in ctx.timings().items():
                sm[name] = sm.get(name, 0.0) + ms / steps
        b.record()
        torch.cuda.synchronize()
        ctx.enable_timing(False)
        ms_ = a.elapsed_time(b) / steps
        if dist is not None:
            t = torch.tensor([ms_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        z, c = batch.extract_sizes(), batch.count_sizes()
        ext_b = t_ + 1.25 * z.hoco_bases + 20.0 * z.n_syncmers
        ext_ms = sm["encode"] + sm["scan"] + sm["kmerhash"]
        return {"k": int(K if k is None else k), "s": S, "reads_per_gpu": int(n_), "value": world * t_ / (ms_ * 1e-3), "unit": "bases/s",
                "ms_per_step": ms_, "steps": steps, "warmup": 3, "extract_ms": ext_ms, "extract_gbs": ext_b / (ext_ms * 1e-3) / 1e9,
                "stage_ms": sm, "syncmers_rank0": int(z.n_syncmers), "distinct_kmers_rank0": int(c.n_unique)}

    # configs[4]: the same reads at k = 501 and k = 2001 (every N the driver launches carries the sweep)
    sweep = None
    if not args.no_sweep and args.k == 1001 and args.workload == "uniform":
        sweep = [timed(kk, None, max(2, min(args.steps, 3))) for kk in (501, 2001)]

    # configs[3]: ONE read set of 10 M x 15 kb reads in contiguous blocks over 8 GPUs, global ids checked on a sample
    config3 = None
    if not args.no_config3 and args.workload == "uniform" and ((world == 8 and not args.config3_reads) or (world > 1 and args.config3_reads)):
        n3 = (args.config3_reads or 10_000_000) // world
        b3, o3 = synth_gpu.hifi_reads_gpu(3000 + rank, 500_000_000, n3, READ_LEN, ERR, dev, genome_seed=3)
        batch.set_sid_base(rank * n3)
        config3 = timed(None, (b3, o3, n3, n3 * READ_LEN), 2)
        config3["workload"] = "BASELINE.json configs[3]: %d x %d b reads of a 500 Mb genome, %d per GPU in contiguous blocks, tuple exchange + id return over NCCL" % (
            n3 * world, READ_LEN, n3)
        config3["global_ids_sample_check"] = check_global_ids(torch, dist, batch, comm, rank, world, dev, n3)
        batch.set_sid_base(rank * n_reads)
        del b3, o3

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    if sweep:
        for e_ in sweep:
            e_["roofline_frac"] = e_["extract_gbs"] / peak
    if config3:
        config3["roofline_frac"] = config3["extract_gbs"] / peak
    # algorithmic bytes of the extract kernels: raw read + hoco_s + ho_rl written + 20 B per syncmer
    ext_bytes = total + 1.25 * sizes.hoco_bases + 20.0 * sizes.n_syncmers
    ext_ms = (stage_ms["encode"] + stage_ms["scan"] + stage_ms["kmerhash"]) / args.steps
    achieved = ext_bytes / (ext_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "extract = encode_kernel + scan_kernel + kmerhash_kernel",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "algorithmic_bytes_per_launch": ext_bytes,
                "bytes_per_raw_base": ext_bytes / total, "ms_per_launch": ext_ms,
                "stage_ms": {k_: v / args.steps for k_, v in stage_ms.items()},
                "stage_launches": {k_: v // args.steps for k_, v in stage_launch.items()},
                "stage_ms_by_rank": by_rank,
                "note": "integer-issue bound, not HBM bound: see DESIGN.md section 5"}

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ns = min(10000 * threads, 160000, n_reads)
        sample = bases[:ns * READ_LEN].cpu().numpy()
        r = cpu_reference_run(sample, ns, threads)
        if r is not None:
            cpu = {"value": r["bases"] / r["seconds"], "unit": "bases/s", "cores": threads, "kind": "reference",
                   "sample": "first %d reads of the step's batch (%.2f Gbases), reference sr_read -t %d + sr_db_stat + collect, FASTA in /dev/shm" % (
                       ns, ns * READ_LEN / 1e9, threads), "stages_s": r["stages"]}

    scan_info = int(batch.debug_scan_info())
    whole = None
    if world == 1 and not args.no_whole:
        del bases, off
        batch.close()
        torch.cuda.empty_cache()
        try:
            whole = whole_command(torch, synth_gpu, dev, args.whole_reads, os.cpu_count() or 1)
        except Exception as e:                         # reported, never fatal for the headline
            whole = {"error": repr(e)[:300]}

    line = {"metric": "HiFi bases/sec syncmer-extract+count", "value": value, "unit": "bases/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(world, n_reads, args.workload, n_planted),
            "clocks": clocks, "numa": numa, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "k_sweep": sweep, "config3": config3, "whole_command": whole,
            "multi_gpu_parity": mgpu, "global_ids_sample_check": ids_check,
            "exchange": None if comm is None else {"transport": "sg_comm_* in C: grouped ncclSend/ncclRecv, counts on the device",
                                                   "bytes_sent_per_step_rank0": ex_bytes},
            "results": {"syncmers": int(sizes.n_syncmers), "distinct_kmers": int(csz.n_unique), "hoco_bases": int(sizes.hoco_bases),
                        "reads_on_exact_scan_path": scan_info}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def whole_command(torch, synth_gpu, dev, n_reads, threads):
    """BASELINE.json configs[2]: the WHOLE command, FASTA -> <out>.utg.gfa + <out>.utg.final.gfa, with the reference's defaults
    (-k 1001 -s 31 -c 30 -a 0.35, read error correction, 3 unzip rounds): once through this repository's syncasm()
    (oatk_b200/host/liboatk_gpu.so over libsyncgpu.so, second call in the process timed) and its `syncasm` command (one
    shot, process start and CUDA context included), once through the unmodified reference's syncasm() on the host cores
    (oracle/_ref/libref.so -- the checker and the CPU baseline). Both output files must be byte-identical."""
    import hashlib
    import numpy as np
    G = 10_000_000
    bases, _ = synth_gpu.hifi_reads_gpu(2, G, n_reads, READ_LEN, ERR, dev, genome_seed=2)
    h = bases.cpu().numpy()
    del bases
    torch.cuda.empty_cache()
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(d, "reads.fa")
    write_fasta(fa, h, n_reads, READ_LEN)
    del h
    out = {"workload": "BASELINE.json configs[2]: syncasm -k 1001 -s 31 -c 30 -a 0.35 -t %d on %d x %d b reads of a %d b genome (%.2f Gbases), FASTA in /dev/shm" % (
        threads, n_reads, READ_LEN, G, n_reads * READ_LEN / 1e9), "raw_bases": n_reads * READ_LEN}
    try:
        from oatk_b200.host import build_host
        H = C.CDLL(build_host.build())
        proto = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                 C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]

        def call(L, prefix):
            L.syncasm.restype = C.c_int
            L.syncasm.argtypes = proto
            files = (C.c_char_p * 1)(fa.encode())
            t0 = time.perf_counter()
            rc = L.syncasm(files, 1, 0, 1001, 31, 100000, 10000, 30, 0.35, 0.3, 1, 3, threads, prefix.encode(), None, 0)
            return rc, time.perf_counter() - t0

        def digest(prefix):
            r = {}
            for sfx in (".utg.gfa", ".utg.final.gfa"):
                x = open(prefix + sfx, "rb").read()
                r[sfx] = {"md5": hashlib.md5(x).hexdigest(), "bytes": len(x), "S": x.count(b"\nS\t"), "L": x.count(b"\nL\t")}
            return r

        po, pc, pr = os.path.join(d, "ours"), os.path.join(d, "ours_cli"), os.path.join(d, "ref")
        sys.stderr.flush()
        rc, t_first = call(H, po)
        assert rc == 0, "syncasm() returned %d" % rc
        rc, t_ours = call(H, po)
        assert rc == 0
        out["ours_s"], out["ours_first_call_s"] = t_ours, t_first
        out["ours"] = digest(po)
        exe = os.path.join(ROOT, "oatk_b200", "host", "syncasm")
        if os.path.exists(exe):
            t0 = time.perf_counter()
            rc = subprocess.call([exe, "-k", "1001", "-s", "31", "-c", "30", "-t", str(threads), "-o", pc, fa],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out["ours_cli_one_shot_s"] = time.perf_counter() - t0
            out["ours_cli_identical"] = rc == 0 and digest(pc) == out["ours"]
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        if pyoracle.have_ref():
            R = pyoracle.Ref().L
            rc, t_ref = call(R, pr)
            out["reference_s"], out["reference_threads"] = t_ref, threads
            out["reference"] = digest(pr) if rc == 0 else None
            out["identical"] = rc == 0 and out["reference"] == out["ours"]
            out["speedup"] = t_ref / t_ours
        else:
            out["reference"] = "oracle/_ref/libref.so not built"
        out["gbases_per_s"] = n_reads * READ_LEN / t_ours / 1e9
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    return out


def share_unique_id(torch, dist, lib, rank, dev):
    """rank 0's NCCL unique id (sg_comm_unique_id) to every rank over the launcher's process group"""
    raw = lib.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def check_global_ids(torch, dist, batch, comm, rank, world, dev, n_reads, sample=4096):
    """configs[3]: global ids at full size. For the syncmers of every rank's first reads: (i) over all ranks, sorting the
    sampled (hash, id) pairs by hash gives non-decreasing ids, equal exactly when the hashes are equal (ids are ranks in
    hash order, syncmer.c:1419-1446); (ii) every sampled id equals id_base(owner) + the position of the hash in the
    owner's sorted table of distinct hashes -- looked up in the owner's table directly, not through the return path."""
    from oatk_b200 import dist as sgdist
    import numpy as np
    kp, kn = batch.buffer("key")
    m = int(min(sample, kn))
    keys = sgdist.tensor_from_ptr(kp, kn, dev)[:m].clone()
    f = batch.extract_download(want_seq=False)
    ids = torch.from_numpy((f["k_mer"][:m] >> np.uint64(1)).astype(np.int64)).to(dev)
    allk = [torch.empty(m, dtype=torch.int64, device=dev) for _ in range(world)]
    alli = [torch.empty(m, dtype=torch.int64, device=dev) for _ in range(world)]
    ms = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ms, torch.tensor([m], dtype=torch.int64, device=dev))
    if len({int(x.item()) for x in ms}) != 1:
        return {"ok": False, "why": "ranks hold different sample sizes"}
    dist.all_gather(allk, keys)
    dist.all_gather(alli, ids)
    k_all = torch.cat(allk).cpu().numpy().view(np.uint64)
    i_all = torch.cat(alli).cpu().numpy()
    order = np.argsort(k_all, kind="stable")
    ks, is_ = k_all[order], i_all[order]
    mono = bool(np.all(np.diff(is_) >= 0) and np.array_equal(np.diff(is_) == 0, np.diff(ks) == 0))
    # owner look-up: this rank answers for the sampled hashes that fall into its range
    hp, hn = batch.buffer("scm_h")
    table = sgdist.tensor_from_ptr(hp, hn, dev).cpu().numpy().view(np.uint64)
    uniq = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(uniq, torch.tensor([hn], dtype=torch.int64, device=dev))
    base = sum(int(u.item()) for u in uniq[:rank])
    pos = np.searchsorted(table, k_all)
    mine = (pos < len(table)) & (table[np.minimum(pos, max(len(table) - 1, 0))] == k_all) if len(table) else np.zeros(len(k_all), bool)
    want = np.where(mine, base + pos, 0).astype(np.int64)
    t = torch.from_numpy(np.stack([want, mine.astype(np.int64)])).to(dev)
    dist.all_reduce(t)
    want_all, owners = t[0].cpu().numpy(), t[1].cpu().numpy()
    owned_once = bool(np.all(owners == 1))
    exact = bool(np.array_equal(want_all, i_all))
    return {"ok": mono and owned_once and exact, "sampled_syncmers": int(len(k_all)), "ids_follow_hash_order": mono,
            "every_hash_has_one_owner": owned_once, "ids_equal_base_plus_rank_in_owner_table": exact,
            "distinct_kmers_total": int(sum(int(u.item()) for u in uniq))}


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("extract_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def run_e2e(torch, lib, ctx, batch, bases, off, n_reads, total, sizes, csz, steps, dist, dev, world, rank=0, full=False):
    """host buffers in, host buffers out through the C ABI: sg_pipe_run_host (chunked over 3 streams:
    upload, kernels and download overlap) -> sg_stat -> sg_count -> sg_count_download on the master batch.
    Every input byte crosses PCIe inside the timed region and every result array lands in pinned host memory.
    full=False (the headline, what syncasm() does): the run lengths (ho_rl, 1 B per hoco base, whose one consumer -- the
    run-length consensus -- is served on the device by sg_runlen_sums) and the packed bases (hoco_s, 1/4 B per hoco base:
    read error correction and the consensus texts are served by sg_ec_correct / sg_kmer_codes) stay in HBM; the syncmer
    lists of every read, the syncmer database and the ids are downloaded.
    full=True: both travel too, i.e. every field of the reference's sr_t lands on the host (the round-1 e2e)."""
    pin = lambda n, dt: torch.empty(max(int(n), 1), dtype=dt, pin_memory=True)
    # pinned host memory: ~31 KB per read (input + every output array). All ranks of a node share the
    # host RAM, so the e2e batch is capped to what fits in half of what is available now.
    full_reads = n_reads
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    fit = int(0.5 * avail / max(world, 1) / 33000)
    if fit < n_reads:
        n_reads = max(1024, fit)
        total = n_reads * READ_LEN
        scale = n_reads / full_reads
        sizes = _Scaled(sizes, scale * 1.05)
        csz = _Scaled(csz, min(1.0, scale * 1.3))
    h_bases = pin(total, torch.uint8)
    h_bases.copy_(bases[:total])
    h_off = pin(n_reads + 1, torch.int64)
    h_off.copy_(off[:n_reads + 1])
    N, U = int(sizes.n_syncmers), int(csz.n_unique)
    slack = 1.02 if world == 1 else 1.15          # a hash range holds about, not exactly, 1/world of the tuples
    o = lib.ExtractOut()
    bufs = {
        "hoco_l": pin(n_reads, torch.int32), "n_scm": pin(n_reads, torch.int32),
        "hoco_s_off": pin(n_reads + 1, torch.int64), "ho_rl_off": pin(n_reads + 1, torch.int64), "scm_off": pin(n_reads + 1, torch.int64),
        "hoco_s_buf": pin(sizes.hoco_s_bytes * slack + 4096 if full else 16, torch.uint8),
        "ho_rl_buf": pin(sizes.ho_rl_bytes * slack + 4096 if full else 16, torch.uint8),
        "m_pos": pin(N * slack, torch.int32), "s_mer": pin(N * slack, torch.int64), "k_mer": pin(N * slack, torch.int64),
        "amb_sid": pin(1024, torch.int32), "amb_pos": pin(1024, torch.int32),
        "lrl_sid": pin(1024, torch.int32), "lrl_idx": pin(1024, torch.int32), "lrl_val": pin(1024, torch.int32),
    }
    for name, _ in lib.ExtractOut._fields_:
        setattr(o, name, bufs[name].data_ptr())
    o.k_mer = None     # k_mer[] is delivered once, as ids, by sg_count_download (the reference overwrites the hashes too, syncmer.c:1378)
    if not full:
        o.ho_rl_buf = None                         # stays on the device (sg_pipe keeps it in the master batch)
        o.hoco_s_buf = None                        # and so do the packed bases: their consumers are served there (sg_kmer_codes, sg_ec_correct)
    caps = lib.PipeCaps(int(N * slack), bufs["hoco_s_buf"].numel(), bufs["ho_rl_buf"].numel(), 1024, 1024)
    co = lib.CountOut()
    cb = {"h": pin(U * slack, torch.int64), "s": pin(U * slack, torch.int64), "cov": pin(U * slack, torch.int32),
          "occ_off": pin(U * slack + 1, torch.int64), "occ": pin(N * slack, torch.int64), "k_mer_id": pin(N * slack, torch.int64)}
    for name, _ in lib.CountOut._fields_:
        setattr(co, name, cb[name].data_ptr())
    L = lib.library()
    n_slots = env_int("SG_PIPE_SLOTS", 6)
    pipe = lib.Pipe(ctx.device, n_slots)
    chunk = env_int("SG_PIPE_CHUNK", 4096)
    comm2 = None
    ko = None
    if world > 1:
        pipe.set_sid_base(rank * full_reads)
        comm2 = lib.Comm(pipe.ctx, world, rank, share_unique_id(torch, dist, lib, rank, dev))
        ko = lib.ExtractOut()                      # the reads' k_mer[] as global ids: one more download of 8 B per syncmer
        ko.k_mer = bufs["k_mer"].data_ptr()
    torch.cuda.synchronize()

    def one():
        z = pipe.run_host(h_bases.data_ptr(), h_off.data_ptr(), n_reads, K, S, chunk, o, caps)
        if comm2 is not None:
            comm2.exchange_tuples(pipe.master)
        pipe.master.stat()
        pipe.master.count()
        if comm2 is not None:
            comm2.return_ids(pipe.master)
            lib._ck(pipe.ctx.h, L.sg_extract_download(pipe.master.h, C.byref(ko)), "sg_extract_download")
        lib._ck(pipe.ctx.h, L.sg_count_download(pipe.master.h, C.byref(co)), "sg_count_download")
        return z

    one()
    one()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = pipe.launches()
    t0 = time.perf_counter()
    for _ in range(steps):
        z = one()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    launches = (pipe.launches() - l0) // steps
    if dist is not None:
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    N = int(z.n_syncmers)
    U = int(pipe.master.count_sizes().n_unique)
    h2d = total + 8 * (n_reads + n_reads // chunk + 1)
    NA = int(pipe.master.count_sizes().n_syncmers)          # tuples of this rank's hash range (== N on one GPU)
    d2h = int((z.hoco_s_bytes + z.ho_rl_bytes if full else 0) + 12 * N + 8 * n_reads + 28 * U + 16 * NA + (8 * N if world > 1 else 0))
    res = {"value": world * total / sec, "unit": "bases/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
           "ms_per_step": sec * 1e3, "steps": steps, "reads_per_gpu": n_reads, "gpu_launches_per_step": int(launches), "chunk_reads": chunk, "streams": n_slots,
           "run_lengths": "downloaded with the packed bases (every sr_t field on the host)" if full else "resident in HBM with the packed bases, served by sg_runlen_sums / sg_kmer_codes / sg_ec_correct (what syncasm() does)",
           "path": "sg_pipe_run_host (chunks of %d reads over several streams) -> sg_stat -> sg_count -> sg_count_download; pinned host buffers; "
                   "%s" % (chunk, "one GPU" if world == 1 else "with the tuple exchange and the id return over NCCL (sg_comm_*) between extract and count; "
                                  "each rank downloads its hash range of the database and the global ids of its reads")}
    if comm2 is not None:
        comm2.close()
    pipe.close()
    res["pcie"] = pcie_probe(torch, dev, h_bases)
    return res


class _Scaled:
    """size estimates for a smaller e2e batch"""

    def __init__(self, src, f):
        for name in ("n_syncmers", "hoco_s_bytes", "ho_rl_bytes", "n_unique"):
            if hasattr(src, name):
                setattr(self, name, int(getattr(src, name) * f) + 4096)


def pcie_probe(torch, dev, h_src):
    """what the host link of this box delivers (pinned memory, 1 GiB copies): the ceiling of the e2e number"""
    n = min(h_src.numel(), 1 << 30)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h_dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d.copy_(h_src[:n], non_blocking=True)

    d2 = torch.empty(n, dtype=torch.uint8, device=dev)

    def d2h():
        with torch.cuda.stream(s2):
            h_dst.copy_(d2, non_blocking=True)

    def both():
        h2d()
        d2h()

    t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
    return {"h2d_gbs": n / t_h2d / 1e9, "d2h_gbs": n / t_d2h / 1e9, "duplex_each_gbs": n / t_both / 1e9,
            "note": "e2e moves ~1 B/base up and ~1 B/base down, so its ceiling is duplex_each_gbs bases/ns"}


if __name__ == "__main__":
    sys.exit(main())
