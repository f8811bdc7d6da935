"""Minimal BAM/BGZF reader and writer (pure Python).  TEST INFRASTRUCTURE: the independent implementation the native I/O library
(telr_b200/csrc/telr_io.cpp) is checked against, and the writer of the test BAMs.  The product reads BAMs through telr_io only.

The reference collects, per locus, the names of all BAM records overlapping the +-1 kb breakpoint window
with ``pysam.AlignmentFile.fetch`` (TELR_assembly.py:385-408).  pysam/htslib are not available in this
image, so this module restates what is needed: BGZF is a series of gzip members (Python's gzip module reads
them transparently), BAM records carry refID/pos/CIGAR from which the reference span is derived.  ``fetch``
has pysam's semantics: every record (primary, secondary, supplementary alike) whose reference interval
[pos, end) overlaps the half-open query interval; records with no reference span count as length 1.
The writer emits valid BGZF blocks + EOF marker so files made by tests are readable by htslib tools.
"""
from __future__ import annotations

import bisect
import gzip
import struct
import zlib

_CONSUMES_REF = {0, 2, 3, 7, 8}      # M D N = X


class BamIndex:
    def __init__(self, path: str):
        with gzip.open(path, "rb") as fh:
            data = fh.read()
        if data[:4] != b"BAM\x01":
            raise ValueError(f"{path}: not a BAM file")
        (l_text,) = struct.unpack_from("<i", data, 4)
        off = 8 + l_text
        (n_ref,) = struct.unpack_from("<i", data, off)
        off += 4
        self.refs = []
        for _ in range(n_ref):
            (l_name,) = struct.unpack_from("<i", data, off)
            name = data[off + 4: off + 4 + l_name - 1].decode()
            (l_ref,) = struct.unpack_from("<i", data, off + 4 + l_name)
            self.refs.append((name, l_ref))
            off += 8 + l_name
        self.by_ref = {name: [] for name, _ in self.refs}
        n = len(data)
        while off + 4 <= n:
            (block_size,) = struct.unpack_from("<i", data, off)
            rec = off + 4
            ref_id, pos, l_read_name, _mapq, _bin, n_cigar, _flag, _l_seq = struct.unpack_from("<iiBBHHHi", data, rec)
            name = data[rec + 32: rec + 32 + l_read_name - 1].decode()
            cig_off = rec + 32 + l_read_name
            span = 0
            for k in range(n_cigar):
                (c,) = struct.unpack_from("<I", data, cig_off + 4 * k)
                if (c & 0xF) in _CONSUMES_REF:
                    span += c >> 4
            if ref_id >= 0 and pos >= 0:
                self.by_ref[self.refs[ref_id][0]].append((pos, pos + max(span, 1), name))
            off = rec + block_size
        self._starts = {}
        self._maxspan = {}
        for r, lst in self.by_ref.items():
            lst.sort(key=lambda t: t[0])
            self._starts[r] = [t[0] for t in lst]
            self._maxspan[r] = max((t[1] - t[0] for t in lst), default=0)

    def fetch(self, chrom: str, start: int, end: int):
        """Yield names of records overlapping [start, end) on chrom, in coordinate order."""
        if chrom not in self.by_ref:
            raise ValueError(f"invalid contig `{chrom}`")      # pysam raises ValueError here
        lst, starts = self.by_ref[chrom], self._starts[chrom]
        lo = bisect.bisect_left(starts, start - self._maxspan[chrom])
        hi = bisect.bisect_left(starts, end)
        for i in range(lo, hi):
            p, e, name = lst[i]
            if e > start and p < end:
                yield name


def _bgzf_block(payload: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(payload) + comp.flush()
    bsize = len(body) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + body + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bam(path: str, refs, records):
    """refs: [(name, length)]; records: iterable of dicts(name, ref_id, pos, flag, cigar=[(op,len)], seq_len)."""
    out = bytearray()
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    out += b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        nb = n.encode() + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    for r in records:
        nb = r["name"].encode() + b"\0"
        cig = r.get("cigar", [])
        l_seq = r.get("seq_len", 0)
        body = struct.pack("<iiBBHHHiiii", r["ref_id"], r["pos"], len(nb), r.get("mapq", 60), 4680, len(cig), r.get("flag", 0),
                           l_seq, -1, -1, 0)
        body += nb + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cig)
        body += bytes((l_seq + 1) // 2) + b"\xff" * l_seq
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        for i in range(0, len(out), 60000):
            fh.write(_bgzf_block(bytes(out[i:i + 60000])))
        fh.write(_BGZF_EOF)
