"""Packed stage-4 batch (host side) and its ctypes view of ``telr_af_batch`` (include/telr_af.h).

A batch is what TELR's get_af() holds after ``prep_assembly_inputs(read_type="all")``
(reference TELR_assembly.py:384-462) and the contig / annotation lookups (TELR_te.py:607-675):
for every locus its reads, its polished contig and the TE interval on it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

PRESETS = {"map-ont": 0, "map-pb": 1, "map-hifi": 2}


class CBatch(C.Structure):
    _fields_ = [
        ("preset", C.c_int32), ("flank_len", C.c_int32), ("flank_off", C.c_int32),
        ("te_len", C.c_int32), ("te_off", C.c_int32), ("n_loci", C.c_int32), ("n_reads", C.c_int32),
        ("n_bases", C.c_int64),
        ("seq2", C.c_void_p), ("nmask", C.c_void_p), ("read_off", C.c_void_p), ("read_len", C.c_void_p),
        ("read_hash", C.c_void_p), ("locus_read_begin", C.c_void_p), ("contig_off", C.c_void_p),
        ("contig_len", C.c_void_p), ("te_start", C.c_void_p), ("te_end", C.c_void_p),
    ]


class CAln(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "mlen", "blen", "n_cigar")] + \
               [("cigar_off", C.c_int64)] + [(n, C.c_int32) for n in ("mapq", "dp_score", "cnt", "score", "subsc", "n_ambi", "inv", "n_sub")]


ALN_DTYPE = np.dtype([(n, "<i4") for n in
                      ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "mlen", "blen", "n_cigar")]
                     + [("cigar_off", "<i8")] + [(n, "<i4") for n in ("mapq", "dp_score", "cnt", "score", "subsc", "n_ambi", "inv", "n_sub")])
assert ALN_DTYPE.itemsize == C.sizeof(CAln)


class CResult(C.Structure):
    _fields_ = [
        ("cov2x", C.c_void_p), ("af", C.c_void_p), ("depth", C.c_void_p),
        ("aln", C.c_void_p), ("aln_cap", C.c_int64), ("cigar", C.c_void_p), ("cigar_cap", C.c_int64),
        ("n_aln", C.c_int64), ("n_cigar", C.c_int64),
        ("dp_cells", C.c_int64), ("n_minimizers", C.c_int64), ("n_anchors", C.c_int64),
        ("n_dp_tasks", C.c_int64), ("n_aln_blocks", C.c_int64),
        ("ms_stage", C.c_float * 8),
    ]


@dataclass
class Batch:
    preset: int
    seq2: np.ndarray            # uint32 [n_bases/16]
    nmask: np.ndarray           # uint32 [n_bases/32]
    read_off: np.ndarray        # int64  [n_reads]
    read_len: np.ndarray        # int32  [n_reads]
    read_hash: np.ndarray       # uint32 [n_reads]
    locus_read_begin: np.ndarray  # int32 [n_loci+1]
    contig_off: np.ndarray      # int64  [n_loci]
    contig_len: np.ndarray      # int32  [n_loci]
    te_start: np.ndarray        # int32  [n_loci]
    te_end: np.ndarray          # int32  [n_loci]
    flank_len: int = 100
    flank_off: int = 200
    te_len: int = 50
    te_off: int = 50
    meta: dict = field(default_factory=dict)

    @property
    def n_loci(self) -> int:
        return int(self.contig_len.shape[0])

    @property
    def n_reads(self) -> int:
        return int(self.read_len.shape[0])

    @property
    def n_bases(self) -> int:
        return int(self.seq2.shape[0]) * 16

    def h2d_bytes(self) -> int:
        return sum(int(a.nbytes) for a in (self.seq2, self.nmask, self.read_off, self.read_len, self.read_hash,
                                           self.locus_read_begin, self.contig_off, self.contig_len,
                                           self.te_start, self.te_end))

    def validate(self) -> None:
        assert self.seq2.dtype == np.uint32 and self.nmask.dtype == np.uint32
        assert self.seq2.shape[0] % 4 == 0 and self.nmask.shape[0] * 2 == self.seq2.shape[0]
        assert self.read_off.dtype == np.int64 and self.contig_off.dtype == np.int64
        for a in (self.read_len, self.locus_read_begin, self.contig_len, self.te_start, self.te_end):
            assert a.dtype == np.int32
        assert self.read_hash.dtype == np.uint32
        assert self.locus_read_begin.shape[0] == self.n_loci + 1
        assert int(self.locus_read_begin[-1]) == self.n_reads
        assert (self.read_off % 64 == 0).all() and (self.contig_off % 64 == 0).all()

    def as_c(self) -> CBatch:
        """ctypes struct over the numpy buffers (the arrays must outlive the struct)."""
        def p(a):
            assert a.flags["C_CONTIGUOUS"]
            return a.ctypes.data
        return CBatch(self.preset, self.flank_len, self.flank_off, self.te_len, self.te_off,
                      self.n_loci, self.n_reads, self.n_bases,
                      p(self.seq2), p(self.nmask), p(self.read_off), p(self.read_len), p(self.read_hash),
                      p(self.locus_read_begin), p(self.contig_off), p(self.contig_len),
                      p(self.te_start), p(self.te_end))

    # ---- helpers used by tests and the host stage ----
    def unpack(self, off: int, length: int) -> np.ndarray:
        """nt4 codes (0..3, 4 = N) of the sequence at base offset ``off``."""
        idx = np.arange(off, off + length, dtype=np.int64)
        c = (self.seq2[idx >> 4] >> ((idx & 15) * 2).astype(np.uint32)) & 3
        n = (self.nmask[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1
        return np.where(n == 1, 4, c).astype(np.uint8)

    def slice(self, l0: int, l1: int) -> "Batch":
        """Loci [l0, l1) as VIEWS of this batch's packed sequences (no re-packing): valid when the batch is packed locus by locus
        (every sequence of a locus lies between the first sequence of that locus and the first of the next), as the native
        gather and the synthetic generator pack it.  The small per-read / per-locus arrays are rebased copies."""
        l0, l1 = int(l0), int(l1)
        r0, r1 = int(self.locus_read_begin[l0]), int(self.locus_read_begin[l1])
        starts = []
        if l1 > l0:
            starts.append(int(self.contig_off[l0:l1].min()))
        if r1 > r0:
            starts.append(int(self.read_off[r0:r1].min()))
        if not starts:
            z = np.zeros(0, np.uint32)
            return Batch(self.preset, z, z, np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint32), np.zeros(1, np.int32),
                         np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                         self.flank_len, self.flank_off, self.te_len, self.te_off, {"owner": self})
        beg = min(starts)
        ends = [int((self.contig_off[l0:l1] + (self.contig_len[l0:l1].astype(np.int64) + 63) // 64 * 64).max())]
        if r1 > r0:
            ends.append(int((self.read_off[r0:r1] + (self.read_len[r0:r1].astype(np.int64) + 63) // 64 * 64).max()))
        end = max(ends)
        assert beg % 64 == 0 and end % 64 == 0
        return Batch(self.preset, self.seq2[beg // 16: end // 16], self.nmask[beg // 32: end // 32],
                     self.read_off[r0:r1] - beg, self.read_len[r0:r1].copy(), self.read_hash[r0:r1].copy(),
                     (self.locus_read_begin[l0:l1 + 1] - r0).astype(np.int32), self.contig_off[l0:l1] - beg, self.contig_len[l0:l1].copy(),
                     self.te_start[l0:l1].copy(), self.te_end[l0:l1].copy(),
                     self.flank_len, self.flank_off, self.te_len, self.te_off, {"owner": self})

    def is_packed_by_locus(self) -> bool:
        """True when loci occupy disjoint, ascending base ranges (what `slice` needs)."""
        if self.n_loci == 0:
            return True
        lrb = self.locus_read_begin
        first = self.contig_off.copy()
        has = lrb[1:] > lrb[:-1]
        idx = np.nonzero(has)[0]
        if len(idx):
            rmin = np.minimum.reduceat(self.read_off, lrb[:-1][has]) if self.n_reads else np.zeros(0, np.int64)
            first[idx] = np.minimum(first[idx], rmin)
            rmax = np.maximum.reduceat(self.read_off, lrb[:-1][has])
            last = self.contig_off.copy()
            last[idx] = np.maximum(last[idx], rmax)
        else:
            last = self.contig_off.copy()
        return bool((first[1:] > last[:-1]).all())

    def subset(self, loci) -> "Batch":
        """A new batch holding only ``loci`` (in that order); sequences are re-packed contiguously."""
        loci = list(loci)
        seqs, roff, rlen, rhash, lrb, coff, clen, ts, te = [], [], [], [], [0], [], [], [], []
        off = 0
        chunks2, chunksn = [], []

        def take(o, ln):
            nonlocal off
            nb = (ln + 63) // 64 * 64
            chunks2.append(self.seq2[o // 16:(o + nb) // 16])
            chunksn.append(self.nmask[o // 32:(o + nb) // 32])
            r = off
            off += nb
            return r
        for l in loci:
            coff.append(take(int(self.contig_off[l]), int(self.contig_len[l])))
            clen.append(int(self.contig_len[l])); ts.append(int(self.te_start[l])); te.append(int(self.te_end[l]))
            for r in range(int(self.locus_read_begin[l]), int(self.locus_read_begin[l + 1])):
                roff.append(take(int(self.read_off[r]), int(self.read_len[r])))
                rlen.append(int(self.read_len[r])); rhash.append(int(self.read_hash[r]))
            lrb.append(len(rlen))
        z2 = np.concatenate(chunks2) if chunks2 else np.zeros(0, np.uint32)
        zn = np.concatenate(chunksn) if chunksn else np.zeros(0, np.uint32)
        return Batch(self.preset, np.ascontiguousarray(z2), np.ascontiguousarray(zn),
                     np.array(roff, np.int64), np.array(rlen, np.int32), np.array(rhash, np.uint32),
                     np.array(lrb, np.int32), np.array(coff, np.int64), np.array(clen, np.int32),
                     np.array(ts, np.int32), np.array(te, np.int32),
                     self.flank_len, self.flank_off, self.te_len, self.te_off, dict(self.meta))


def pack_sequences(seqs, lib=None):
    """Pack ASCII (bytes/str) sequences; returns (seq2, nmask, offsets[int64], lengths[int32])."""
    lens = np.array([len(s) for s in seqs], np.int32)
    padded = (lens.astype(np.int64) + 63) // 64 * 64
    offs = np.zeros(len(seqs), np.int64)
    if len(seqs):
        offs[1:] = np.cumsum(padded)[:-1]
    nb = int(padded.sum())
    seq2 = np.zeros(nb // 16, np.uint32)
    nmask = np.zeros(nb // 32, np.uint32)
    lut = np.full(256, 4, np.uint8)
    for ch, v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
        lut[ord(ch)] = v
        lut[ord(ch.lower())] = v
    for s, o in zip(seqs, offs):
        b = s.encode() if isinstance(s, str) else bytes(s)
        if lib is not None:
            rc = lib.telr_pack_seq(b, len(b), int(o), seq2.ctypes.data, nmask.ctypes.data)
            if rc != 0:
                raise RuntimeError("telr_pack_seq failed")
            continue
        c = lut[np.frombuffer(b, np.uint8)]
        idx = np.arange(int(o), int(o) + len(b), dtype=np.int64)
        code = np.where(c < 4, c, 0).astype(np.uint32)
        np.bitwise_or.at(seq2, idx >> 4, code << ((idx & 15) * 2).astype(np.uint32))
        nm = (c == 4).astype(np.uint32)
        np.bitwise_or.at(nmask, idx >> 5, nm << (idx & 31).astype(np.uint32))
    return seq2, nmask, offs, lens


def name_hash(name: str) -> int:
    """minimap2 __ac_X31_hash_string of the read name (feeds region tie-breaks)."""
    b = name.encode()
    if not b:
        return 0
    h = b[0]
    for ch in b[1:]:
        h = ((h << 5) - h + ch) & 0xFFFFFFFF
    return h
