"""ctypes binding of the host I/O library (include/telr_io.h -> telr_b200/_telr_io.so).

Read gather for stage 4 (reference: prep_assembly_inputs(read_type="all") + extract_reads, TELR_assembly.py:384-471):
indexed BAM window queries, one streaming pass over the raw reads, packing straight into the telr_af_batch layout; plus
`samtools index` and the sorted-BAM writer used for the `-k` intermediates (TELR_te.py:507-512).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_telr_io.so")
EXPORTS = ["telr_io_strerror", "telr_bam_open", "telr_bam_close", "telr_bam_n_ref", "telr_bam_ref_name", "telr_bam_ref_len", "telr_bam_tid",
           "telr_bam_fetch", "telr_bam_blocks_inflated", "telr_bam_index_build", "telr_gather_run", "telr_gather_free", "telr_bam_write_sorted"]
_LIB = None


class IoError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().telr_io_strerror(code).decode()
        super().__init__(f"{detail}: {msg}" if detail else msg)


class GatherIn(C.Structure):
    _fields_ = [("bam_path", C.c_char_p), ("raw_reads_path", C.c_char_p), ("n_loci", C.c_int32),
                ("chrom", C.POINTER(C.c_char_p)), ("win_beg", C.POINTER(C.c_int64)), ("win_end", C.POINTER(C.c_int64)),
                ("contig_seq", C.POINTER(C.c_char_p)), ("contig_len", C.POINTER(C.c_int32)),
                ("reads_dir", C.c_char_p), ("locus_name", C.POINTER(C.c_char_p)), ("n_threads", C.c_int32)]


class GatherOut(C.Structure):
    _fields_ = [("n_live", C.c_int32), ("n_reads", C.c_int32), ("n_bases", C.c_int64),
                ("seq2", C.POINTER(C.c_uint32)), ("nmask", C.POINTER(C.c_uint32)),
                ("read_off", C.POINTER(C.c_int64)), ("read_len", C.POINTER(C.c_int32)), ("read_hash", C.POINTER(C.c_uint32)),
                ("locus_read_begin", C.POINTER(C.c_int32)), ("contig_off", C.POINTER(C.c_int64)), ("contig_len", C.POINTER(C.c_int32)),
                ("live_index", C.POINTER(C.c_int32)), ("n_names", C.POINTER(C.c_int32)),
                ("t_bam_s", C.c_double), ("t_reads_s", C.c_double), ("t_pack_s", C.c_double), ("t_write_s", C.c_double),
                ("reads_scanned", C.c_int64), ("bases_scanned", C.c_int64), ("unique_reads", C.c_int64), ("bgzf_blocks", C.c_int64),
                ("err", C.c_char * 256)]


class SamRec(C.Structure):
    _fields_ = [("qname", C.c_char_p), ("flag", C.c_int32), ("tid", C.c_int32), ("pos", C.c_int32), ("mapq", C.c_int32),
                ("cigar", C.c_void_p), ("n_cigar", C.c_int32), ("seq", C.c_char_p), ("l_seq", C.c_int32),
                ("aux", C.c_char_p), ("l_aux", C.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(SO_PATH)
        L.telr_io_strerror.restype = C.c_char_p
        L.telr_io_strerror.argtypes = [C.c_int]
        L.telr_bam_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.telr_bam_close.argtypes = [C.c_void_p]
        L.telr_bam_n_ref.argtypes = [C.c_void_p]
        L.telr_bam_ref_name.restype = C.c_char_p
        L.telr_bam_ref_name.argtypes = [C.c_void_p, C.c_int]
        L.telr_bam_ref_len.restype = C.c_int64
        L.telr_bam_ref_len.argtypes = [C.c_void_p, C.c_int]
        L.telr_bam_tid.argtypes = [C.c_void_p, C.c_char_p]
        L.telr_bam_fetch.restype = C.c_int64
        L.telr_bam_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.telr_bam_blocks_inflated.restype = C.c_int64
        L.telr_bam_blocks_inflated.argtypes = [C.c_void_p]
        L.telr_bam_index_build.argtypes = [C.c_char_p, C.c_char_p]
        L.telr_gather_run.argtypes = [C.POINTER(GatherIn), C.POINTER(GatherOut)]
        L.telr_gather_free.argtypes = [C.POINTER(GatherOut)]
        L.telr_bam_write_sorted.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_char_p, C.c_int64,
                                            C.POINTER(SamRec), C.c_int32]
        _LIB = L
    return _LIB


class BamFile:
    """Indexed BAM: `fetch(chrom, start, end)` with pysam's semantics (TELR_assembly.py:385-408), served from the .bai."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        rc = lib().telr_bam_open(path.encode(), C.byref(self._h))
        if rc != 0:
            raise IoError(rc, path)
        n = lib().telr_bam_n_ref(self._h)
        self.refs = [(lib().telr_bam_ref_name(self._h, i).decode(), int(lib().telr_bam_ref_len(self._h, i))) for i in range(n)]

    def fetch(self, chrom: str, start: int, end: int):
        tid = lib().telr_bam_tid(self._h, chrom.encode())
        if tid < 0:
            raise ValueError(f"invalid contig `{chrom}`")          # pysam raises ValueError here
        p, nb = C.c_void_p(), C.c_int64()
        n = lib().telr_bam_fetch(self._h, tid, int(start), int(end), C.byref(p), C.byref(nb))
        if n < 0:
            raise IoError(int(n), "fetch")
        if n == 0:
            return []
        return [s.decode() for s in C.string_at(p, nb.value).split(b"\0")[:-1]]

    @property
    def blocks_inflated(self) -> int:
        return int(lib().telr_bam_blocks_inflated(self._h))

    def close(self):
        if self._h:
            lib().telr_bam_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:     # noqa: BLE001
            pass


def index_bam(bam_path: str, bai_path: str | None = None):
    """`samtools index`."""
    rc = lib().telr_bam_index_build(bam_path.encode(), (bai_path or bam_path + ".bai").encode())
    if rc != 0:
        raise IoError(rc, bam_path)


class Gathered:
    """Result of `gather`: numpy views of the packed batch arrays (owned by this object) + what the gather cost."""

    def __init__(self, out: GatherOut, n_loci: int):
        self._out = out
        o = out

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).view(dt) if n else np.zeros(0, dt)
        self.n_live, self.n_reads, self.n_bases = o.n_live, o.n_reads, o.n_bases
        self.seq2 = arr(o.seq2, o.n_bases // 16, np.uint32)
        self.nmask = arr(o.nmask, o.n_bases // 32, np.uint32)
        self.read_off = arr(o.read_off, o.n_reads, np.int64)
        self.read_len = arr(o.read_len, o.n_reads, np.int32)
        self.read_hash = arr(o.read_hash, o.n_reads, np.uint32)
        self.locus_read_begin = arr(o.locus_read_begin, o.n_live + 1, np.int32)
        self.contig_off = arr(o.contig_off, o.n_live, np.int64)
        self.contig_len = arr(o.contig_len, o.n_live, np.int32)
        self.live_index = arr(o.live_index, o.n_live, np.int32).copy()
        self.n_names = arr(o.n_names, n_loci, np.int32).copy()
        self.timing = dict(bam_s=o.t_bam_s, reads_s=o.t_reads_s, pack_s=o.t_pack_s, write_s=o.t_write_s, reads_scanned=o.reads_scanned,
                           bases_scanned=o.bases_scanned, unique_reads=o.unique_reads, bgzf_blocks=o.bgzf_blocks)

    def free(self):
        if self._out is not None:
            lib().telr_gather_free(C.byref(self._out))
            self._out = None

    def __del__(self):
        try:
            self.free()
        except Exception:     # noqa: BLE001
            pass


def gather(bam: str, raw_reads: str, chroms, win_beg, win_end, contigs, reads_dir=None, locus_names=None, threads: int = 0) -> Gathered:
    """Window queries + one pass over the raw reads + packing.  `contigs[l]` is the locus's contig (bytes) or None."""
    n = len(chroms)
    gin = GatherIn()
    gin.bam_path, gin.raw_reads_path, gin.n_loci = bam.encode(), raw_reads.encode(), n
    ch = (C.c_char_p * max(n, 1))(*[c.encode() for c in chroms])
    wb = np.ascontiguousarray(win_beg, np.int64)
    we = np.ascontiguousarray(win_end, np.int64)
    cs = (C.c_char_p * max(n, 1))(*[(c if c else None) for c in contigs])
    cl = np.array([len(c) if c else 0 for c in contigs], np.int32)
    ln = (C.c_char_p * max(n, 1))(*[(s.encode() if locus_names else None) for s in (locus_names or [""] * n)])
    gin.chrom, gin.contig_seq, gin.locus_name = ch, cs, ln
    gin.win_beg = wb.ctypes.data_as(C.POINTER(C.c_int64))
    gin.win_end = we.ctypes.data_as(C.POINTER(C.c_int64))
    gin.contig_len = cl.ctypes.data_as(C.POINTER(C.c_int32))
    gin.reads_dir = reads_dir.encode() if reads_dir else None
    gin.n_threads = threads
    out = GatherOut()
    rc = lib().telr_gather_run(C.byref(gin), C.byref(out))
    if rc != 0:
        detail = out.err.decode(errors="replace")
        if rc == -6:
            raise ValueError(detail)                 # pysam: invalid contig
        if rc == -7:
            raise KeyError(detail)                   # SeqIO.index(...).get_raw
        raise IoError(rc, detail)
    return Gathered(out, n)
