"""ctypes binding of libsyncgpu.so (include/syncgpu.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if
that fails, or no CUDA device is present when a compute call is made, the call
raises. Nothing here imports or calls oracle/.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OATK_SYNCGPU_LIB") or os.path.join(HERE, "libsyncgpu.so")   # the override is for kernel experiments (tools/)

SG_T_NAMES = ["encode", "scan", "kmerhash", "place", "sort", "group", "stat", "arcs", "pack", "ec", "exchange", "ids"]


class SgError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__("%s failed: %s (%d)%s" % (what, _lib().sg_strerror(code).decode(), code,
                                                    (": " + detail) if detail else ""))


class ExtractSizes(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_reads", "n_syncmers", "hoco_bases", "hoco_s_bytes", "ho_rl_bytes", "n_ambiguous", "n_long_runs")]


class ExtractOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("hoco_l", "n_scm", "hoco_s_off", "ho_rl_off", "scm_off", "hoco_s_buf", "ho_rl_buf",
                 "m_pos", "s_mer", "k_mer", "amb_sid", "amb_pos", "lrl_sid", "lrl_idx", "lrl_val")]


class StatOut(C.Structure):
    _fields_ = [("n_syncmers", C.c_uint64), ("n_gaps", C.c_uint64), ("gap_sum", C.c_int64),
                ("smer_unique", C.c_uint64), ("smer_singleton", C.c_uint64),
                ("kmer_unique", C.c_uint64), ("kmer_singleton", C.c_uint64),
                ("smer_cnts", C.c_int64 * 1001), ("kmer_cnts", C.c_int64 * 1001)]


class CountSizes(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_syncmers", "n_unique", "n_hash_collisions")]


class PipeCaps(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("max_syncmers", "hoco_s_bytes", "ho_rl_bytes", "max_ambiguous", "max_long_runs")]


class CountOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("h", "s", "cov", "occ_off", "occ", "k_mer_id")]


_LIB = None

# every symbol include/syncgpu.h declares; tests check that the library exports all of them
SYMBOLS = [
    "sg_ctx_create", "sg_host_bind_near_device", "sg_ctx_destroy", "sg_ctx_set_stream", "sg_ctx_sync", "sg_strerror", "sg_last_error",
    "sg_ctx_launches", "sg_ctx_enable_timing", "sg_ctx_timings",
    "sg_batch_create", "sg_batch_destroy", "sg_batch_set_reads_host", "sg_batch_set_reads_device",
    "sg_batch_set_sid_base", "sg_extract", "sg_extract_sizes", "sg_extract_download",
    "sg_stat", "sg_stat_multiplicities", "sg_count", "sg_count_conflict", "sg_count_sizes", "sg_count_download", "sg_arcs", "sg_arcs_download",
    "sg_tuples_partition", "sg_tuples_adopt", "sg_debug_set_hash_bits", "sg_debug_set_sort_low_bits", "sg_debug_sort_info", "sg_debug_set_pack_bits", "sg_debug_scan_info", "sg_batch_buffer",
    "sg_ids_pack", "sg_ids_scatter", "sg_batch_set_exact_verify", "sg_smer_counts_pack", "sg_smer_counts_merge", "sg_batch_set_lists_host",
    "sg_comm_unique_id", "sg_comm_init_rank", "sg_comm_init_all", "sg_comm_destroy", "sg_comm_rank", "sg_comm_world", "sg_comm_bytes_sent",
    "sg_comm_exchange_tuples", "sg_comm_global_stat", "sg_comm_return_ids", "sg_comm_arcs",
    "sg_pipe_create", "sg_pipe_destroy", "sg_pipe_run_host", "sg_pipe_run_host_cb", "sg_pipe_master", "sg_pipe_ctx", "sg_pipe_last_error", "sg_pipe_launches", "sg_pipe_set_sid_base", "sg_pipe_keep_run_lengths", "sg_pipe_keep_packed_bases", "sg_kmer_codes", "sg_pipe_set_capacity_factor", "sg_pipe_syncmer_overflow", "sg_runlen_sums", "sg_runlen_resident",
    "sg_ec_correct", "sg_ec_result_free", "sg_debug_set_ec_arena", "sg_ec_filter", "sg_ec_filter_free", "sg_arc_votes",
]


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        from . import build as _b
        _b.build()
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.sg_strerror.restype = C.c_char_p
    L.sg_strerror.argtypes = [i32]
    L.sg_last_error.restype = C.c_char_p
    L.sg_last_error.argtypes = [vp]
    L.sg_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.sg_host_bind_near_device.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.sg_ctx_destroy.argtypes = [vp]
    L.sg_ctx_destroy.restype = None
    L.sg_ctx_set_stream.argtypes = [vp, vp]
    L.sg_ctx_sync.argtypes = [vp]
    L.sg_ctx_launches.restype = u64
    L.sg_ctx_launches.argtypes = [vp]
    L.sg_ctx_enable_timing.argtypes = [vp, i32]
    L.sg_ctx_timings.argtypes = [vp, vp, vp]
    L.sg_batch_create.argtypes = [vp, C.POINTER(vp)]
    L.sg_batch_destroy.argtypes = [vp]
    L.sg_batch_destroy.restype = None
    L.sg_batch_set_reads_host.argtypes = [vp, vp, vp, u64]
    L.sg_batch_set_reads_device.argtypes = [vp, vp, vp, u64, u64]
    L.sg_batch_set_sid_base.argtypes = [vp, u64]
    L.sg_extract.argtypes = [vp, i32, i32]
    L.sg_extract_sizes.argtypes = [vp, C.POINTER(ExtractSizes)]
    L.sg_extract_download.argtypes = [vp, C.POINTER(ExtractOut)]
    for name, args in (("sg_stat", [vp, C.POINTER(StatOut)]), ("sg_count", [vp]), ("sg_count_conflict", [vp, C.POINTER(u64)]),
                       ("sg_count_sizes", [vp, C.POINTER(CountSizes)]), ("sg_count_download", [vp, C.POINTER(CountOut)]),
                       ("sg_arcs", [vp, C.c_uint32, C.c_double, C.POINTER(u64)]), ("sg_arcs_download", [vp, vp]),
                       ("sg_tuples_partition", [vp, i32, vp, C.POINTER(vp)]), ("sg_tuples_adopt", [vp, vp, u64]),
                       ("sg_debug_set_hash_bits", [vp, i32]), ("sg_debug_set_sort_low_bits", [vp, i32]),
                       ("sg_debug_sort_info", [vp, C.POINTER(u64), C.POINTER(i32)]), ("sg_debug_set_pack_bits", [vp, i32]), ("sg_debug_scan_info", [vp, C.POINTER(u64)]), ("sg_batch_buffer", [vp, i32, C.POINTER(vp), C.POINTER(u64)]),
                       ("sg_ids_pack", [vp, u64, C.POINTER(vp), C.POINTER(u64)]), ("sg_smer_counts_pack", [vp, C.POINTER(vp), C.POINTER(u64)]),
                       ("sg_smer_counts_merge", [vp, vp, u64, vp]), ("sg_ids_scatter", [vp, vp, u64]), ("sg_batch_set_exact_verify", [vp, i32])):
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    L.sg_comm_unique_id.argtypes = [vp]
    L.sg_comm_init_rank.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.sg_comm_init_all.argtypes = [C.POINTER(vp), i32, C.POINTER(vp)]
    L.sg_comm_destroy.argtypes = [vp]
    L.sg_comm_destroy.restype = None
    L.sg_comm_bytes_sent.argtypes = [vp]
    L.sg_comm_bytes_sent.restype = u64
    L.sg_comm_exchange_tuples.argtypes = [vp, vp]
    L.sg_comm_global_stat.argtypes = [vp, vp, vp]
    L.sg_comm_return_ids.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(u64)]
    L.sg_comm_arcs.argtypes = [vp, vp, C.c_uint32, C.c_double, i32, C.POINTER(u64)]
    L.sg_pipe_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.sg_pipe_destroy.argtypes = [vp]
    L.sg_pipe_destroy.restype = None
    L.sg_pipe_run_host.argtypes = [vp, vp, vp, u64, i32, i32, u64, C.POINTER(ExtractOut), C.POINTER(PipeCaps), C.POINTER(ExtractSizes)]
    L.sg_pipe_master.restype = vp
    L.sg_pipe_master.argtypes = [vp]
    L.sg_pipe_ctx.restype = vp
    L.sg_pipe_ctx.argtypes = [vp]
    L.sg_pipe_last_error.restype = C.c_char_p
    L.sg_pipe_last_error.argtypes = [vp]
    L.sg_pipe_launches.restype = u64
    L.sg_pipe_launches.argtypes = [vp]
    _LIB = L
    return L


def library():
    return _lib()


def bind_host_near_device(device, cpus=True, memory=True):
    """sg_host_bind_near_device: CPUs and / or NUMA node next to the GPU for this process; returns (numa_node, n_cpus)."""
    node, ncpu = C.c_int32(-1), C.c_int32(0)
    rc = _lib().sg_host_bind_near_device(int(device), (1 if cpus else 0) | (2 if memory else 0), C.byref(node), C.byref(ncpu))
    if rc != 0:
        raise SgError(rc, "sg_host_bind_near_device", "")
    return node.value, ncpu.value


def _ck(ctx, rc, what):
    if rc != 0:
        detail = _lib().sg_last_error(ctx).decode() if ctx else ""
        raise SgError(rc, what, detail)


class Context:
    """one per GPU (reference has no analogue: it is the device the path runs on)"""

    def __init__(self, device=0, stream=None):
        L = _lib()
        h = C.c_void_p()
        rc = L.sg_ctx_create(device, C.byref(h))
        if rc != 0:
            raise SgError(rc, "sg_ctx_create", "no usable CUDA device %d; libsyncgpu has no CPU fallback" % device)
        self.h = h
        self.device = device
        if stream is not None:
            _ck(self.h, L.sg_ctx_set_stream(self.h, C.c_void_p(stream)), "sg_ctx_set_stream")

    def sync(self):
        _ck(self.h, _lib().sg_ctx_sync(self.h), "sg_ctx_sync")

    def launches(self):
        return int(_lib().sg_ctx_launches(self.h))

    def enable_timing(self, on=True):
        _ck(self.h, _lib().sg_ctx_enable_timing(self.h, 1 if on else 0), "sg_ctx_enable_timing")

    def timings(self):
        ms = (C.c_float * len(SG_T_NAMES))()
        ln = (C.c_uint32 * len(SG_T_NAMES))()
        _ck(self.h, _lib().sg_ctx_timings(self.h, ms, ln), "sg_ctx_timings")
        return {n: (float(ms[i]), int(ln[i])) for i, n in enumerate(SG_T_NAMES)}

    def close(self):
        if self.h:
            _lib().sg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        _ck(ctx.h, _lib().sg_batch_create(ctx.h, C.byref(h)), "sg_batch_create")
        self.h = h
        self._keep = None

    def set_reads_host(self, bases, off):
        """bases: uint8 numpy array (all reads back to back); off: uint64 offsets, n+1"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self._keep = (bases, off)
        _ck(self.ctx.h, _lib().sg_batch_set_reads_host(self.h, bases.ctypes.data, off.ctypes.data, len(off) - 1),
            "sg_batch_set_reads_host")

    def set_reads_host_ptr(self, bases_ptr, off_ptr, n_reads):
        _ck(self.ctx.h, _lib().sg_batch_set_reads_host(self.h, bases_ptr, off_ptr, n_reads), "sg_batch_set_reads_host")

    def set_reads_device(self, d_bases_ptr, d_off_ptr, n_reads, total_bases):
        _ck(self.ctx.h, _lib().sg_batch_set_reads_device(self.h, d_bases_ptr, d_off_ptr, n_reads, total_bases),
            "sg_batch_set_reads_device")

    def set_sid_base(self, sid_base):
        _ck(self.ctx.h, _lib().sg_batch_set_sid_base(self.h, sid_base), "sg_batch_set_sid_base")

    def extract(self, k, s):
        _ck(self.ctx.h, _lib().sg_extract(self.h, k, s), "sg_extract")

    def extract_sizes(self):
        z = ExtractSizes()
        _ck(self.ctx.h, _lib().sg_extract_sizes(self.h, C.byref(z)), "sg_extract_sizes")
        return z

    def extract_download(self, want_seq=True):
        """returns one flat dictionary of unpadded arrays in read order (the layout the parity tests compare)"""
        z = self.extract_sizes()
        n, N = z.n_reads, z.n_syncmers
        a = dict(
            hoco_l=np.zeros(n, np.uint32), n_scm=np.zeros(n, np.uint32),
            hoco_s_off=np.zeros(n + 1, np.uint64), ho_rl_off=np.zeros(n + 1, np.uint64), scm_off=np.zeros(n + 1, np.uint64),
            hoco_s_buf=np.zeros(z.hoco_s_bytes + 16, np.uint8) if want_seq else None,
            ho_rl_buf=np.zeros(z.ho_rl_bytes + 16, np.uint8) if want_seq else None,
            m_pos=np.zeros(N, np.uint32), s_mer=np.zeros(N, np.uint64), k_mer=np.zeros(N, np.uint64),
            amb_sid=np.zeros(z.n_ambiguous, np.uint32), amb_pos=np.zeros(z.n_ambiguous, np.uint32),
            lrl_sid=np.zeros(z.n_long_runs, np.uint32), lrl_idx=np.zeros(z.n_long_runs, np.uint32),
            lrl_val=np.zeros(z.n_long_runs, np.uint32))
        o = ExtractOut()
        for name, _t in ExtractOut._fields_:
            arr = a[name]
            setattr(o, name, arr.ctypes.data if arr is not None and arr.size else None)
        _ck(self.ctx.h, _lib().sg_extract_download(self.h, C.byref(o)), "sg_extract_download")
        f = dict(hoco_l=a["hoco_l"], n_scm=a["n_scm"], m_pos=a["m_pos"], s_mer=a["s_mer"], k_mer=a["k_mer"],
                 n_lrl=np.bincount(a["lrl_sid"], minlength=n).astype(np.uint32) if n else np.zeros(0, np.uint32),
                 n_n=np.bincount(a["amb_sid"], minlength=n).astype(np.uint32) if n else np.zeros(0, np.uint32),
                 ho_l_rl=a["lrl_val"], n_nucl=a["amb_pos"], lrl_idx=a["lrl_idx"])
        if want_seq:
            hl = a["hoco_l"].astype(np.int64)
            hsb = (hl + 3) // 4
            f["hoco_s"] = _unpad(a["hoco_s_buf"], a["hoco_s_off"], hsb)
            f["ho_rl"] = _unpad(a["ho_rl_buf"], a["ho_rl_off"], hl)
        return f

    def stat(self):
        st = StatOut()
        _ck(self.ctx.h, _lib().sg_stat(self.h, C.byref(st)), "sg_stat")
        return st

    def count(self):
        _ck(self.ctx.h, _lib().sg_count(self.h), "sg_count")

    def count_conflict(self):
        """after count() raised SG_E_SMER_CONFLICT: (k-mer hash, s-mer code 0, read 0, s-mer code 1, read 1)"""
        out = (C.c_uint64 * 5)()
        _ck(self.ctx.h, _lib().sg_count_conflict(self.h, out), "sg_count_conflict")
        return tuple(int(x) for x in out)

    def count_sizes(self):
        z = CountSizes()
        _ck(self.ctx.h, _lib().sg_count_sizes(self.h, C.byref(z)), "sg_count_sizes")
        return z

    def count_download(self):
        z = self.count_sizes()
        U, N = z.n_unique, z.n_syncmers
        a = dict(h=np.zeros(U, np.uint64), s=np.zeros(U, np.uint64), cov=np.zeros(U, np.uint32),
                 occ_off=np.zeros(U + 1, np.uint64), occ=np.zeros(N, np.uint64), k_mer_id=np.zeros(N, np.uint64))
        o = CountOut()
        for name, _t in CountOut._fields_:
            setattr(o, name, a[name].ctypes.data if a[name].size else None)
        _ck(self.ctx.h, _lib().sg_count_download(self.h, C.byref(o)), "sg_count_download")
        a["off"] = a.pop("occ_off")
        a["n_hash_collisions"] = int(z.n_hash_collisions)
        return a

    def arcs(self, min_k_cov, min_a_cov_f):
        n = C.c_uint64(0)
        _ck(self.ctx.h, _lib().sg_arcs(self.h, min_k_cov, min_a_cov_f, C.byref(n)), "sg_arcs")
        out = np.zeros((n.value, 4), np.uint64)
        if n.value:
            _ck(self.ctx.h, _lib().sg_arcs_download(self.h, out.ctypes.data), "sg_arcs_download")
        return out

    BUF = dict(key=0, occ=1, smer=2, mpos=3, kid=4, sorted_occ=5, scm_h=6, scm_cov=7, adopted_occ=8)

    def buffer(self, which):
        """(device pointer, element count) of one of the batch's device arrays"""
        p, n = C.c_void_p(), C.c_uint64()
        _ck(self.ctx.h, _lib().sg_batch_buffer(self.h, self.BUF[which], C.byref(p), C.byref(n)), "sg_batch_buffer")
        return p.value or 0, n.value

    def tuples_partition(self, n_parts):
        counts = (C.c_uint64 * n_parts)()
        p = C.c_void_p()
        _ck(self.ctx.h, _lib().sg_tuples_partition(self.h, n_parts, counts, C.byref(p)), "sg_tuples_partition")
        return [int(x) for x in counts], p.value or 0

    def tuples_adopt(self, d_ptr, n):
        _ck(self.ctx.h, _lib().sg_tuples_adopt(self.h, d_ptr, n), "sg_tuples_adopt")

    def ids_pack(self, id_base):
        p, n = C.c_void_p(), C.c_uint64()
        _ck(self.ctx.h, _lib().sg_ids_pack(self.h, id_base, C.byref(p), C.byref(n)), "sg_ids_pack")
        return p.value or 0, n.value

    def ids_scatter(self, d_ptr, n):
        _ck(self.ctx.h, _lib().sg_ids_scatter(self.h, d_ptr, n), "sg_ids_scatter")

    def set_exact_verify(self, on=True):
        _ck(self.ctx.h, _lib().sg_batch_set_exact_verify(self.h, 1 if on else 0), "sg_batch_set_exact_verify")

    def debug_set_hash_bits(self, bits):
        _ck(self.ctx.h, _lib().sg_debug_set_hash_bits(self.h, bits), "sg_debug_set_hash_bits")

    def smer_counts_pack(self):
        p, n = C.c_void_p(), C.c_uint64(0)
        _ck(self.ctx.h, _lib().sg_smer_counts_pack(self.h, C.byref(p), C.byref(n)), "sg_smer_counts_pack")
        return p.value, int(n.value)

    def smer_counts_merge(self, d_ptr, n, stat):
        _ck(self.ctx.h, _lib().sg_smer_counts_merge(self.h, d_ptr, n, C.byref(stat)), "sg_smer_counts_merge")
        return stat

    def debug_set_sort_low_bits(self, bits):
        _ck(self.ctx.h, _lib().sg_debug_set_sort_low_bits(self.h, bits), "sg_debug_set_sort_low_bits")

    def debug_set_pack_bits(self, bits):
        _ck(self.ctx.h, _lib().sg_debug_set_pack_bits(self.h, bits), "sg_debug_set_pack_bits")

    def debug_sort_info(self):
        r, f = C.c_uint64(0), C.c_int(0)
        _ck(self.ctx.h, _lib().sg_debug_sort_info(self.h, C.byref(r), C.byref(f)), "sg_debug_sort_info")
        return int(r.value), bool(f.value)

    def runlen_sums(self, occ_off, occ, k):
        """occ_off: n+1 offsets, occ: read << 32 | start << 1 | strand; returns (n, k) uint64 sums of run length - 1"""
        occ_off = np.ascontiguousarray(occ_off, dtype=np.uint64)
        occ = np.ascontiguousarray(occ, dtype=np.uint64)
        n = len(occ_off) - 1
        out = np.zeros((n, k), np.uint64)
        L = _lib()
        L.sg_runlen_sums.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _ck(self.ctx.h, L.sg_runlen_sums(self.h, n, occ_off.ctypes.data, occ.ctypes.data if len(occ) else None, out.ctypes.data), "sg_runlen_sums")
        return out

    def debug_scan_info(self):
        r = C.c_uint64(0)
        _ck(self.ctx.h, _lib().sg_debug_scan_info(self.h, C.byref(r)), "sg_debug_scan_info")
        return int(r.value)

    def close(self):
        if self.h:
            _lib().sg_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128 bytes that rank 0 hands to the other ranks (ncclGetUniqueId)"""
    buf = C.create_string_buffer(128)
    rc = _lib().sg_comm_unique_id(buf)
    if rc:
        raise SgError("sg_comm_unique_id: %s" % _lib().sg_strerror(rc).decode())
    return buf.raw


class Comm:
    """the multi-GPU exchange in C over NCCL (csrc/sg_comm.cu); one per rank, every call is collective"""

    def __init__(self, ctx, world, rank, unique_id):
        self.ctx, self.world, self.rank = ctx, world, rank
        h = C.c_void_p()
        _ck(ctx.h, _lib().sg_comm_init_rank(ctx.h, world, rank, unique_id, C.byref(h)), "sg_comm_init_rank")
        self.h = h

    def exchange_tuples(self, batch):
        _ck(self.ctx.h, _lib().sg_comm_exchange_tuples(self.h, batch.h), "sg_comm_exchange_tuples")

    def global_stat(self, batch, st):
        _ck(self.ctx.h, _lib().sg_comm_global_stat(self.h, batch.h, C.byref(st)), "sg_comm_global_stat")
        return st

    def return_ids(self, batch):
        base, tot = C.c_uint64(0), C.c_uint64(0)
        _ck(self.ctx.h, _lib().sg_comm_return_ids(self.h, batch.h, C.byref(base), C.byref(tot)), "sg_comm_return_ids")
        return int(base.value), int(tot.value)

    def arcs(self, batch, min_k_cov, min_a_cov_f, root=0):
        n = C.c_uint64(0)
        _ck(self.ctx.h, _lib().sg_comm_arcs(self.h, batch.h, min_k_cov, min_a_cov_f, root, C.byref(n)), "sg_comm_arcs")
        out = np.zeros((n.value, 4), np.uint64)
        if n.value:
            _ck(self.ctx.h, _lib().sg_arcs_download(batch.h, out.ctypes.data), "sg_arcs_download")
        return out

    def bytes_sent(self):
        return int(_lib().sg_comm_bytes_sent(self.h))

    def close(self):
        if self.h:
            _lib().sg_comm_destroy(self.h)
            self.h = None


def _unpad(buf, off, lens):
    """concatenate buf[off[r] : off[r] + lens[r]] over reads"""
    n = len(lens)
    if n == 0:
        return np.zeros(0, np.uint8)
    tot = int(lens.sum())
    out = np.empty(tot, np.uint8)
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]
    # vectorised ragged gather
    idx = np.arange(tot, dtype=np.int64) - np.repeat(starts, lens) + np.repeat(off[:-1].astype(np.int64), lens)
    out[:] = buf[idx]
    return out


class _Borrowed:
    """a handle owned by someone else (the pipeline's master context / batch)"""

    def __init__(self, h, device=0):
        self.h = C.c_void_p(h)
        self.device = device


class Pipe:
    """sg_pipe_*: host buffers in, host buffers out, chunked over several streams"""

    def __init__(self, device=0, n_slots=3):
        h = C.c_void_p()
        rc = _lib().sg_pipe_create(device, n_slots, C.byref(h))
        if rc != 0:
            raise SgError(rc, "sg_pipe_create")
        self.h = h
        self.device = device
        ctx = Context.__new__(Context)
        ctx.h = C.c_void_p(_lib().sg_pipe_ctx(h))
        ctx.device = device
        ctx.close = lambda: None
        self.ctx = ctx
        m = Batch.__new__(Batch)
        m.ctx = ctx
        m.h = C.c_void_p(_lib().sg_pipe_master(h))
        m._keep = None
        m.close = lambda: None
        self.master = m

    def set_sid_base(self, base):
        _lib().sg_pipe_set_sid_base.argtypes = [C.c_void_p, C.c_uint64]
        rc = _lib().sg_pipe_set_sid_base(self.h, base)
        if rc != 0:
            raise SgError(rc, "sg_pipe_set_sid_base")

    def run_host(self, bases_ptr, off_ptr, n_reads, k, s, chunk_reads, out, caps):
        z = ExtractSizes()
        rc = _lib().sg_pipe_run_host(self.h, bases_ptr, off_ptr, n_reads, k, s, chunk_reads, C.byref(out), C.byref(caps), C.byref(z))
        if rc != 0:
            raise SgError(rc, "sg_pipe_run_host", _lib().sg_pipe_last_error(self.h).decode())
        return z

    def launches(self):
        return int(_lib().sg_pipe_launches(self.h))

    def close(self):
        if self.h:
            _lib().sg_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
