"""Execute the REFERENCE's own AF arithmetic (TELR_te.py) by AST extraction.

The reference module cannot be imported (its top-level ``from Bio import SeqIO`` fails here), so
the FunctionDefs get_te_flank_ratio / get_te_cov / get_flank_cov and the three statements of
get_af spanning TELR_te.py:758-835 are compiled from the file where it lies, with
get_median_cov(bam, chr, start, end) stubbed by an in-memory depth array that follows samtools'
region semantics.  Used by tests (when /root/reference exists) and by
tests/golden/make_af_vectors.py to produce committed golden vectors.
"""
from __future__ import annotations

import ast
import os
import statistics

REF = "/root/reference/src/telr/TELR_te.py"


def available() -> bool:
    return os.path.isfile(REF)


def _load():
    src = open(REF).read()
    tree = ast.parse(src)
    fns = {}
    get_af = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            if node.name in ("get_te_flank_ratio", "get_te_cov", "get_flank_cov"):
                fns[node.name] = node
            if node.name == "get_af":
                get_af = node
    # the AF block of get_af: `te_freq = dict()` + the two `with` statements that follow it
    blk = []
    started = False
    for st in get_af.body:
        if (isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and st.targets[0].id == "te_freq"):
            started = True
        if started and isinstance(st, (ast.Assign, ast.With)):
            if isinstance(st, ast.Assign) and st.targets[0].id != "te_freq":
                continue
            blk.append(st)
    assert len(blk) == 3, [type(b) for b in blk]
    return fns, blk


def ref_functions(depth_by_bam):
    """Returns dict of the reference's functions bound to a stub get_median_cov.

    depth_by_bam: {bam_name: list/array of per-base depth}; region "c:S-E" -> D[max(S,1)-1 : min(E, L)]
    """
    fns, _ = _load()

    def get_median_cov(bam, chr, start, end):
        d = depth_by_bam[bam]
        beg = max(int(start) - 1, 0)
        covs = [int(v) for v in d[beg:min(int(end), len(d))]]
        return statistics.median(covs)

    ns = {"get_median_cov": get_median_cov}
    mod = ast.Module(body=list(fns.values()), type_ignores=[])
    exec(compile(mod, REF, "exec"), ns)
    return ns


def ref_af_block(freq_path, freq_rc_path):
    """Run TELR_te.py:758-835 on crafted .freq / .revcomp.freq files; returns te_freq dict."""
    fns, blk = _load()
    ns = {"vcf_parsed_freq": freq_path, "vcf_parsed_freq_revcomp": freq_rc_path}
    mod = ast.Module(body=[fns["get_te_flank_ratio"]] + blk, type_ignores=[])
    exec(compile(mod, REF, "exec"), ns)
    return ns["te_freq"]


def synth_depth(seed: int, L: int, s: int, e: int):
    """Deterministic pseudo-random fw/rc depth arrays (own LCG, so fixtures only need to store the seed)."""
    st = (seed * 2862933555777941757 + 3037000493) & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal st
        st = (st * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return st >> 33
    base = nxt() % 61
    zero = nxt() % 7 == 0
    bump = nxt() % 26
    dfw = [0 if zero else max(0, base + nxt() % 17 - 8 + (15 if s <= i < e else 0)) for i in range(L)]
    drc = [max(0, base + nxt() % 17 - 8 + (bump if L - e <= i < L - s else 0)) for i in range(L)]
    return dfw, drc
