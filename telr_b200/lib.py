"""ctypes binding of the product library telr_b200/_telr_af.so (C ABI: include/telr_af.h).

There is deliberately no CPU fallback: if the CUDA library is missing or no sm_100 device is
present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import ALN_DTYPE, Batch, CBatch, CResult

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_telr_af.so")
_LIB = None

ERRORS = {0: "ok", -1: "EINVAL", -2: "ENOMEM", -3: "ECUDA", -4: "ENODEV", -5: "ECAP", -6: "EUNSUPPORTED"}


class TelrError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().telr_af_strerror(code).decode() if _LIB is not None else ERRORS.get(code, "?")
        super().__init__(f"{where}: {ERRORS.get(code, code)} ({msg})")


class DpTask(C.Structure):
    _fields_ = [("q_off", C.c_int64), ("t_off", C.c_int64), ("qlen", C.c_int32), ("tlen", C.c_int32),
                ("w", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("flag", C.c_int32)]


class DpOut(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score",
                                          "zdropped", "reach_end", "n_cigar")] + [("cigar_off", C.c_int64), ("cells", C.c_int64)]


DPTASK_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"),
                         ("zdrop", "<i4"), ("end_bonus", "<i4"), ("flag", "<i4")])
DPOUT_DTYPE = np.dtype([(n, "<i4") for n in ("max", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score",
                                             "zdropped", "reach_end", "n_cigar")] + [("_pad", "<i4"), ("cigar_off", "<i8"), ("cells", "<i8")])
assert DPTASK_DTYPE.itemsize == C.sizeof(DpTask) and DPOUT_DTYPE.itemsize == C.sizeof(DpOut)

EXPORTS = ["telr_af_create", "telr_af_destroy", "telr_af_run", "telr_af_run_device", "telr_af_sketch", "telr_af_depth_af",
           "telr_af_dp", "telr_af_strerror", "telr_af_last_cuda", "telr_af_version", "telr_af_launch_count", "telr_af_stream",
           "telr_pack_seq", "telr_name_hash", "telr_af_plan_chunks", "telr_af_set_option"]


def lib():
    """Load the CUDA library; raises if it was not built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing - the CUDA extension must be built; there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.telr_af_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t]
        L.telr_af_destroy.argtypes = [C.c_void_p]
        L.telr_af_run.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(CResult)]
        L.telr_af_run_device.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(CResult)]
        L.telr_af_sketch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.telr_af_depth_af.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + \
                                      [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.telr_af_dp.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int64]
        L.telr_af_strerror.restype = C.c_char_p
        L.telr_af_strerror.argtypes = [C.c_int]
        L.telr_af_last_cuda.argtypes = [C.c_void_p]
        L.telr_pack_seq.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        L.telr_name_hash.restype = C.c_uint32
        L.telr_af_launch_count.restype = C.c_longlong
        L.telr_af_launch_count.argtypes = [C.c_void_p]
        L.telr_af_stream.restype = C.c_void_p
        L.telr_af_stream.argtypes = [C.c_void_p]
        L.telr_name_hash.argtypes = [C.c_char_p]
        L.telr_af_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
        L.telr_af_plan_chunks.restype = C.c_int
        L.telr_af_plan_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32]
        for fn in ("telr_af_create", "telr_af_destroy", "telr_af_run", "telr_af_run_device", "telr_af_sketch",
                   "telr_af_depth_af", "telr_af_dp", "telr_af_last_cuda", "telr_af_version", "telr_pack_seq"):
            getattr(L, fn).restype = C.c_int
        _LIB = L
    return _LIB


def plan_chunks(read_len, locus_read_begin, budget_bases):
    """Chunks of loci telr_af_run processes together for a budget of read bases per chunk: [0, ..., n_loci] (host only)."""
    read_len = np.ascontiguousarray(read_len, np.int32)
    lrb = np.ascontiguousarray(locus_read_begin, np.int32)
    n_loci = len(lrb) - 1
    cuts = np.zeros(n_loci + 2, np.int32)
    n = lib().telr_af_plan_chunks(read_len.ctypes.data, lrb.ctypes.data, n_loci, int(budget_bases), cuts.ctypes.data, len(cuts))
    if n < 0:
        raise TelrError(n, "telr_af_plan_chunks")
    return cuts[: n + 1]


class Result:
    """Host-side result buffers of one telr_af_run call."""

    def __init__(self, b: Batch, want_depth=False, want_aln=False, aln_cap=None, cigar_cap=None):
        self.cov2x = np.zeros((b.n_loci, 8), np.int32)
        self.af = np.zeros(b.n_loci, np.float64)
        self.depth = np.zeros(int(2 * b.contig_len.astype(np.int64).sum()), np.int32) if want_depth else None
        self.aln = self.cigar = None
        self.c = CResult()
        self.c.cov2x = self.cov2x.ctypes.data
        self.c.af = self.af.ctypes.data
        if want_depth:
            self.c.depth = self.depth.ctypes.data
        if want_aln:
            aln_cap = aln_cap or (b.n_reads * 2 * 6 + 64)
            cigar_cap = cigar_cap or int(b.read_len.astype(np.int64).sum() * 2 + 4096)
            self.aln = np.zeros(aln_cap, ALN_DTYPE)
            self.cigar = np.zeros(cigar_cap, np.uint32)
            self.c.aln = self.aln.ctypes.data
            self.c.aln_cap = aln_cap
            self.c.cigar = self.cigar.ctypes.data
            self.c.cigar_cap = cigar_cap

    @property
    def alns(self):
        return self.aln[: self.c.n_aln]

    def cigar_of(self, i):
        a = self.aln[i]
        return self.cigar[a["cigar_off"]: a["cigar_off"] + a["n_cigar"]]

    def stats(self) -> dict:
        names = ["sketch", "seed_chain", "plan", "align_dp", "traceback", "finalize", "depth_af", "misc"]
        return {"dp_cells": int(self.c.dp_cells), "n_minimizers": int(self.c.n_minimizers), "n_anchors": int(self.c.n_anchors),
                "n_dp_tasks": int(self.c.n_dp_tasks), "n_aln_blocks": int(self.c.n_aln_blocks),
                "ms": {n: float(self.c.ms_stage[i]) for i, n in enumerate(names)}}


class Context:
    """One telr_af_ctx: owns a device, a stream and its workspace."""

    def __init__(self, device: int = 0, workspace_bytes: int = 0):
        self._h = C.c_void_p()
        rc = lib().telr_af_create(C.byref(self._h), device, workspace_bytes)
        if rc != 0:
            raise TelrError(rc, "telr_af_create")
        self.device = device

    def close(self):
        if self._h:
            lib().telr_af_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(lib().telr_af_launch_count(self._h))

    @property
    def stream_ptr(self) -> int:
        return int(lib().telr_af_stream(self._h) or 0)

    def run(self, b: Batch, **kw) -> Result:
        """Whole stage-4 body on host buffers (H2D + kernels + D2H)."""
        r = Result(b, **kw)
        cb = b.as_c()
        rc = lib().telr_af_run(self._h, C.byref(cb), C.byref(r.c))
        if rc != 0:
            raise TelrError(rc, "telr_af_run")
        return r

    def set_option(self, name: str, value: int):
        """Mapping option on top of the preset, e.g. set_option("bw", 2000) for `minimap2 -r2k`; 0 restores the preset value."""
        rc = lib().telr_af_set_option(self._h, name.encode(), int(value))
        if rc != 0:
            raise TelrError(rc, f"telr_af_set_option({name})")

    def run_device(self, cb: CBatch, cres: CResult):
        rc = lib().telr_af_run_device(self._h, C.byref(cb), C.byref(cres))
        if rc != 0:
            raise TelrError(rc, "telr_af_run_device")

    def sketch(self, seq2, nmask, offs, lens, w, k, hpc=0):
        n_bases = len(seq2) * 16
        cap = int(lens.astype(np.int64).sum()) + 16 * len(lens) + 16
        x = np.zeros(cap, np.uint64)
        y = np.zeros(cap, np.uint64)
        off = np.zeros(len(lens) + 1, np.int64)
        offs = np.ascontiguousarray(offs, np.int64)
        lens = np.ascontiguousarray(lens, np.int32)
        rc = lib().telr_af_sketch(self._h, seq2.ctypes.data, nmask.ctypes.data, n_bases, len(lens), offs.ctypes.data,
                                  lens.ctypes.data, w, k, hpc, x.ctypes.data, y.ctypes.data, cap, off.ctypes.data)
        if rc != 0:
            raise TelrError(rc, "telr_af_sketch")
        return x[: off[-1]], y[: off[-1]], off

    def depth_af(self, contig_len, te_start, te_end, blk_ls, blk_start, blk_len, flank_len=100, flank_off=200,
                 te_len=50, te_off=50):
        n = len(contig_len)
        contig_len = np.ascontiguousarray(contig_len, np.int32)
        te_start = np.ascontiguousarray(te_start, np.int32)
        te_end = np.ascontiguousarray(te_end, np.int32)
        blk_ls = np.ascontiguousarray(blk_ls, np.int32)
        blk_start = np.ascontiguousarray(blk_start, np.int32)
        blk_len = np.ascontiguousarray(blk_len, np.int32)
        depth = np.zeros(int(2 * contig_len.astype(np.int64).sum()), np.int32)
        cov = np.zeros((n, 8), np.int32)
        af = np.zeros(n, np.float64)
        rc = lib().telr_af_depth_af(self._h, n, contig_len.ctypes.data, te_start.ctypes.data, te_end.ctypes.data,
                                    flank_len, flank_off, te_len, te_off, len(blk_ls), blk_ls.ctypes.data,
                                    blk_start.ctypes.data, blk_len.ctypes.data, depth.ctypes.data, cov.ctypes.data,
                                    af.ctypes.data)
        if rc != 0:
            raise TelrError(rc, "telr_af_depth_af")
        return depth, cov, af

    def dp(self, preset, tasks: np.ndarray, qseq: np.ndarray, tseq: np.ndarray, cigar_cap=None):
        tasks = np.ascontiguousarray(tasks, DPTASK_DTYPE)
        out = np.zeros(len(tasks), DPOUT_DTYPE)
        cigar_cap = cigar_cap or int((tasks["qlen"].astype(np.int64) + tasks["tlen"]).sum() + 64)
        cig = np.zeros(cigar_cap, np.uint32)
        qseq = np.ascontiguousarray(qseq, np.uint8)
        tseq = np.ascontiguousarray(tseq, np.uint8)
        rc = lib().telr_af_dp(self._h, preset, len(tasks), tasks.ctypes.data, qseq.ctypes.data, len(qseq), tseq.ctypes.data,
                              len(tseq), out.ctypes.data, cig.ctypes.data, cigar_cap)
        if rc != 0:
            raise TelrError(rc, "telr_af_dp")
        return out, cig
