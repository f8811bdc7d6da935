"""ctypes binding of the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from telr_b200.batch import ALN_DTYPE, Batch, CBatch, CResult

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
_LIB = None


class OrcOpt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "k", "w", "hpc", "a", "b", "q", "e", "q2", "e2", "sc_ambi", "zdrop", "zdrop_inv", "end_bonus",
        "min_dp_max", "min_ksw_len", "bw", "bw_long", "max_gap", "max_chain_skip", "max_chain_iter", "min_cnt",
        "min_chain_score", "rmq_inner_dist", "rmq_size_cap", "rmq_rescue_size")] + [
        ("rmq_rescue_ratio", C.c_float), ("chain_gap_scale", C.c_float), ("chain_skip_scale", C.c_float),
        ("mask_level", C.c_float), ("mask_len", C.c_int32), ("pri_ratio", C.c_float), ("best_n", C.c_int32),
        ("q_occ_frac", C.c_float), ("mid_occ_frac", C.c_float), ("min_mid_occ", C.c_int32),
        ("max_mid_occ", C.c_int32), ("max_max_occ", C.c_int32), ("occ_dist", C.c_int32), ("seed", C.c_int32), ("max_sw_mat", C.c_int64), ("rank_min_len", C.c_int32),
        ("rank_frac", C.c_float), ("max_clip_ratio", C.c_float)]


class OrcEz(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score",
                                          "zdropped", "reach_end", "n_cigar", "m_cigar")] + [
        ("cigar", C.POINTER(C.c_uint32)), ("cells", C.c_int64)]


class OrcDbg(C.Structure):
    _fields_ = [
        ("n_mz", C.c_int64), ("mz_x", C.POINTER(C.c_uint64)), ("mz_y", C.POINTER(C.c_uint64)),
        ("n_a", C.c_int64), ("a", C.POINTER(C.c_uint64)),
        ("n_u", C.c_int32), ("u", C.POINTER(C.c_uint64)), ("n_ca", C.c_int64), ("ca", C.POINTER(C.c_uint64)),
        ("mid_occ", C.c_int32), ("rechained", C.c_int32),
        ("n_regs0", C.c_int32), ("regs0", C.POINTER(C.c_int32)),
    ]


def build() -> str:
    subprocess.run(["make", "-s", "-C", ORC_DIR], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(ORC_DIR, "_build", "liborc.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORC_DIR, "_build", "liborc.so")
        if not os.path.exists(path):
            path = build()
        L = C.CDLL(path)
        L.orc_sketch.restype = C.c_int64
        L.orc_sketch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_af_run.restype = C.c_int
        L.orc_af_run.argtypes = [C.POINTER(CBatch), C.POINTER(CResult), C.c_int, C.c_int, C.c_int]
        L.orc_opt_preset.argtypes = [C.POINTER(OrcOpt), C.c_int]
        L.orc_ksw_extd2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 11 + [C.POINTER(OrcEz)]
        L.orc_ksw_ll.restype = C.c_int
        L.orc_ksw_ll.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.POINTER(C.c_int)] * 2
        L.orc_median2x.restype = C.c_int32
        L.orc_median2x.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_cov_af.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_map_one.restype = C.c_int
        L.orc_map_one.argtypes = [C.POINTER(OrcOpt), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint32,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(OrcDbg)]
        L.orc_dbg_free.argtypes = [C.POINTER(OrcDbg)]
        L.orc_radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_bw.argtypes = [C.c_int, C.c_int]
        L.orc_set_ext_bound.argtypes = [C.c_int]
        L.orc_name_hash.restype = C.c_uint32
        L.orc_name_hash.argtypes = [C.c_char_p]
        _LIB = L
    return _LIB


def opt(preset: int) -> OrcOpt:
    o = OrcOpt()
    lib().orc_opt_preset(C.byref(o), preset)
    return o


def sketch(nt4: np.ndarray, w: int, k: int, hpc: int = 0):
    nt4 = np.ascontiguousarray(nt4, np.uint8)
    cap = len(nt4) + 16
    x = np.zeros(cap, np.uint64)
    y = np.zeros(cap, np.uint64)
    n = lib().orc_sketch(nt4.ctypes.data, len(nt4), w, k, hpc, x.ctypes.data, y.ctypes.data, cap)
    assert n <= cap
    return x[:n].copy(), y[:n].copy()


class Result:
    def __init__(self, b: Batch, want_depth=True, want_aln=True, aln_cap=None, cigar_cap=None):
        self.cov2x = np.zeros((b.n_loci, 8), np.int32)
        self.af = np.zeros(b.n_loci, np.float64)
        self.depth = np.zeros(int(2 * b.contig_len.astype(np.int64).sum()), np.int32) if want_depth else None
        if want_aln:
            aln_cap = aln_cap or (b.n_reads * 2 * 6 + 64)
            cigar_cap = cigar_cap or int(b.read_len.astype(np.int64).sum() * 2 + 4096)
            self.aln = np.zeros(aln_cap, ALN_DTYPE)
            self.cigar = np.zeros(cigar_cap, np.uint32)
        else:
            self.aln = self.cigar = None
        self.c = CResult()
        self.c.cov2x = self.cov2x.ctypes.data
        self.c.af = self.af.ctypes.data
        self.c.depth = self.depth.ctypes.data if want_depth else None
        if want_aln:
            self.c.aln = self.aln.ctypes.data
            self.c.aln_cap = len(self.aln)
            self.c.cigar = self.cigar.ctypes.data
            self.c.cigar_cap = len(self.cigar)

    @property
    def alns(self):
        return self.aln[: self.c.n_aln]

    def cigar_of(self, i):
        a = self.aln[i]
        return self.cigar[a["cigar_off"]: a["cigar_off"] + a["n_cigar"]]


def af_run(b: Batch, threads: int = 0, first: int = 0, n: int = 0, **kw) -> Result:
    r = Result(b, **kw)
    cb = b.as_c()
    rc = lib().orc_af_run(C.byref(cb), C.byref(r.c), threads, first, n)
    if rc != 0:
        raise RuntimeError(f"orc_af_run -> {rc}")
    return r


def ksw_extd2(q: np.ndarray, t: np.ndarray, o: OrcOpt, w: int, zdrop: int, end_bonus: int, flag: int):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    ez = OrcEz()
    lib().orc_ksw_extd2(len(q), q.ctypes.data, len(t), t.ctypes.data, o.a, o.b, o.sc_ambi, o.q, o.e, o.q2, o.e2,
                        w, zdrop, end_bonus, flag, C.byref(ez))
    cig = np.ctypeslib.as_array(ez.cigar, shape=(ez.n_cigar,)).copy() if ez.n_cigar else np.zeros(0, np.uint32)
    d = {n: getattr(ez, n) for n in ("max", "max_q", "max_t", "mqe", "mqe_t", "mte", "score", "zdropped", "reach_end", "cells")}
    d["cigar"] = cig
    if ez.cigar:
        C.CDLL(None).free(ez.cigar)
    return d


def map_one(o: OrcOpt, contig: np.ndarray, read: np.ndarray, name_hash: int = 0, want_dbg=True):
    contig = np.ascontiguousarray(contig, np.uint8)
    read = np.ascontiguousarray(read, np.uint8)
    aln = np.zeros(64, ALN_DTYPE)
    cig = np.zeros(len(read) * 4 + 1024, np.uint32)
    ncig = C.c_int64(0)
    cells = C.c_int64(0)
    tasks = C.c_int64(0)
    dbg = OrcDbg()
    n = lib().orc_map_one(C.byref(o), contig.ctypes.data, len(contig), read.ctypes.data, len(read), name_hash,
                          aln.ctypes.data, len(aln), cig.ctypes.data, len(cig), C.byref(ncig), C.byref(cells),
                          C.byref(tasks), C.byref(dbg) if want_dbg else None)
    out = {"aln": aln[:max(n, 0)].copy(), "cigar": cig[: ncig.value].copy(), "cells": cells.value, "tasks": tasks.value}
    if want_dbg:
        def u64(p, m):
            return np.ctypeslib.as_array(p, shape=(m,)).copy() if m else np.zeros(0, np.uint64)
        out["mz_x"], out["mz_y"] = u64(dbg.mz_x, dbg.n_mz), u64(dbg.mz_y, dbg.n_mz)
        out["anchors"] = u64(dbg.a, 2 * dbg.n_a).reshape(-1, 2)
        out["u"] = u64(dbg.u, dbg.n_u)
        out["chain_anchors"] = u64(dbg.ca, 2 * dbg.n_ca).reshape(-1, 2)
        out["mid_occ"], out["rechained"] = dbg.mid_occ, dbg.rechained
        out["regs0"] = (np.ctypeslib.as_array(dbg.regs0, shape=(dbg.n_regs0 * 10,)).copy().reshape(-1, 10)
                        if dbg.n_regs0 else np.zeros((0, 10), np.int32))
        lib().orc_dbg_free(C.byref(dbg))
    return out


def set_bw(bw: int = 0, bw_long: int = 0):
    """minimap2 -r NUM[,NUM] on top of the preset for af_run (0 = preset value)."""
    lib().orc_set_bw(int(bw), int(bw_long))


def set_ext_bound(on: bool = True):
    """Bounded extension rule of orc_ksw_extd2 (exact work saving); False = every anti-diagonal like ksw2."""
    lib().orc_set_ext_bound(1 if on else 0)
