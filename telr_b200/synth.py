"""Synthetic stage-4 batches: ctypes front end of csrc/telr_synth.c.

Bench and test tooling.  ``CONFIGS`` are the five BASELINE.json configurations (config 1, the repo
test set, is built from /root/reference/test by tests/golden/make_fixture.py and not generated here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import Batch, PRESETS

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SynthCfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_loci_total", C.c_int32), ("depth", C.c_double),
        ("mean_len", C.c_double), ("sigma_len", C.c_double), ("min_len", C.c_int32), ("max_len", C.c_int32),
        ("flank_lo", C.c_int32), ("flank_hi", C.c_int32), ("te_min", C.c_int32), ("te_max", C.c_int32),
        ("te_median", C.c_double), ("te_sigma", C.c_double),
        ("p_sub", C.c_double), ("p_ins", C.c_double), ("p_del", C.c_double), ("hp_mult", C.c_double),
        ("p_polish", C.c_double), ("p_n", C.c_double), ("n_families", C.c_int32),
    ]


class SynthOut(C.Structure):
    _fields_ = [
        ("n_loci", C.c_int32), ("n_reads", C.c_int32), ("n_bases", C.c_int64),
        ("seq2", C.POINTER(C.c_uint32)), ("nmask", C.POINTER(C.c_uint32)),
        ("read_off", C.POINTER(C.c_int64)), ("read_len", C.POINTER(C.c_int32)),
        ("read_hash", C.POINTER(C.c_uint32)), ("locus_read_begin", C.POINTER(C.c_int32)),
        ("contig_off", C.POINTER(C.c_int64)), ("contig_len", C.POINTER(C.c_int32)),
        ("te_start", C.POINTER(C.c_int32)), ("te_end", C.POINTER(C.c_int32)),
        ("truth_af", C.POINTER(C.c_float)), ("read_truth", C.POINTER(C.c_int32)), ("read_origin", C.POINTER(C.c_int32)),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_telr_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
        _LIB.telr_synth_generate.argtypes = [C.POINTER(SynthCfg), C.c_int32, C.c_int32, C.POINTER(SynthOut)]
        _LIB.telr_synth_generate.restype = C.c_int
        _LIB.telr_synth_generate_list.argtypes = [C.POINTER(SynthCfg), C.c_void_p, C.c_int32, C.POINTER(SynthOut)]
        _LIB.telr_synth_generate_list.restype = C.c_int
        _LIB.telr_synth_default.argtypes = [C.POINTER(SynthCfg)]
        _LIB.telr_synth_free.argtypes = [C.POINTER(SynthOut)]
    return _LIB


# name -> (preset, overrides); seeds are 20221101 + config id
CONFIGS = {
    "ont_3k_50x": dict(id=2, preset="map-ont", n_loci=3000, depth=50, mean_len=10000),
    "clr_3k_40x": dict(id=3, preset="map-pb", n_loci=3000, depth=40, mean_len=10000,
                       p_sub=0.015, p_ins=0.08, p_del=0.045),
    "hifi_3k_40x": dict(id=3, preset="map-hifi", n_loci=3000, depth=40, mean_len=12000,
                        p_sub=0.0005, p_ins=0.0003, p_del=0.0002, sigma_len=0.2),
    "ont_30k_30x": dict(id=4, preset="map-ont", n_loci=30000, depth=30, mean_len=10000),
    "poly_10k_200x": dict(id=5, preset="map-ont", n_loci=10000, depth=200, mean_len=10000,
                          te_min=7000, te_max=9000, te_median=8000, flank_lo=3000, flank_hi=3000),
}


def make_cfg(name: str, **over) -> tuple[SynthCfg, str]:
    spec = dict(CONFIGS[name])
    spec.update(over)
    cfg = SynthCfg()
    _lib().telr_synth_default(C.byref(cfg))
    cfg.seed = 20221101 + int(spec.pop("id"))
    preset = spec.pop("preset")
    cfg.n_loci_total = int(spec.pop("n_loci"))
    for k, v in spec.items():
        setattr(cfg, k, v)
    return cfg, preset


def generate(name: str, first_locus: int = 0, n_loci: int | None = None, total_loci: int | None = None, loci=None, **over) -> Batch:
    """Generate loci [first_locus, first_locus+n_loci) of configuration ``name`` (``total_loci`` widens the job), or the
    explicit list ``loci`` of global locus ids (every locus has its own RNG stream, so any subset comes out identical)."""
    cfg, preset = make_cfg(name, **over)
    if total_loci is not None:
        cfg.n_loci_total = int(total_loci)
    out = SynthOut()
    if loci is not None:
        ids = np.ascontiguousarray(loci, np.int32)
        first_locus, n_loci = (int(ids[0]) if len(ids) else 0), len(ids)
        rc = _lib().telr_synth_generate_list(C.byref(cfg), ids.ctypes.data, len(ids), C.byref(out))
    else:
        if n_loci is None:
            n_loci = cfg.n_loci_total - first_locus
        rc = _lib().telr_synth_generate(C.byref(cfg), first_locus, n_loci, C.byref(out))
    if rc != 0:
        raise MemoryError("telr_synth_generate failed")
    try:
        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        nl, nr, nb = out.n_loci, out.n_reads, out.n_bases
        b = Batch(
            PRESETS[preset],
            arr(out.seq2, nb // 16, np.uint32), arr(out.nmask, nb // 32, np.uint32),
            arr(out.read_off, nr, np.int64), arr(out.read_len, nr, np.int32), arr(out.read_hash, nr, np.uint32),
            arr(out.locus_read_begin, nl + 1, np.int32), arr(out.contig_off, nl, np.int64),
            arr(out.contig_len, nl, np.int32), arr(out.te_start, nl, np.int32), arr(out.te_end, nl, np.int32),
            meta={"config": name, "first_locus": first_locus,
                  "truth_af": arr(out.truth_af, nl, np.float32), "read_truth": arr(out.read_truth, nr, np.int32),
                  "read_origin": arr(out.read_origin, 4 * nr, np.int32).reshape(-1, 4)},
        )
    finally:
        _lib().telr_synth_free(C.byref(out))
    b.validate()
    return b
