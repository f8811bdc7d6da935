"""Shared helpers for the test-suite (emu harness binding, comparisons)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_EMU = None


def emu():
    global _EMU
    if _EMU is None:
        L = C.CDLL(os.path.join(ROOT, "tests", "emu", "_emu.so"))
        L.emu_map_from_anchors.restype = C.c_int
        L.emu_map_from_anchors.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int]
        L.emu_radix_sort_128x.argtypes = [C.c_void_p, C.c_int]
        _EMU = L
    return _EMU


def load_config1():
    from telr_b200.batch import Batch
    d = np.load(os.path.join(ROOT, "tests", "golden", "config1_batch.npz"))
    return Batch(int(d["preset"]), d["seq2"], d["nmask"], d["read_off"], d["read_len"], d["read_hash"], d["locus_read_begin"],
                 d["contig_off"], d["contig_len"], d["te_start"], d["te_end"])


ALN_FIELDS = ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "mlen", "blen", "n_cigar",
              "mapq", "dp_score", "cnt", "score", "subsc", "n_ambi", "inv", "n_sub")


def assert_same_results(r, ro, check_aln=True):
    """r: product result, ro: oracle result (both expose cov2x, af, depth, alns, cigar_of)."""
    assert (r.cov2x == ro.cov2x).all(), (r.cov2x[(r.cov2x != ro.cov2x).any(1)][:5], ro.cov2x[(r.cov2x != ro.cov2x).any(1)][:5])
    assert np.array_equal(np.isnan(r.af), np.isnan(ro.af))
    ok = ~np.isnan(ro.af)
    assert np.all(np.abs(r.af[ok] - ro.af[ok]) <= 1e-9 * np.abs(ro.af[ok])), "AF beyond 1e-9 relative"
    if r.depth is not None and ro.depth is not None:
        assert (r.depth == ro.depth).all(), "per-base depth differs"
    if check_aln and r.aln is not None and ro.aln is not None:
        ga, oa = r.alns, ro.alns
        assert len(ga) == len(oa), (len(ga), len(oa))
        for f in ALN_FIELDS:
            assert (ga[f] == oa[f]).all(), (f, np.nonzero(ga[f] != oa[f])[0][:5])
        for i in range(len(ga)):
            assert (r.cigar_of(i) == ro.cigar_of(i)).all(), ("cigar", i)
