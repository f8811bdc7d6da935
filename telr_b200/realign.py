"""SAM/BAM and PAF views of the device alignments: what `minimap2 -a` / `minimap2 -c` print for the same records.

  write_realign_bams   `<prefix>.realign.sort.bam` + `.bai` per locus and contig strand — the intermediates of realignment()
                       (TELR_te.py:495-515: minimap2 -a | samtools view -bS | samtools sort | samtools index), row f2
  align_to_bam         reads vs one contig with a preset and bandwidth, sorted BAM: the polishing alignment
                       `minimap2 -t T -ax <preset> -r2k contig reads | samtools sort` (TELR_assembly.py:199-212), row f1
  align_to_paf         queries vs contigs as PAF with the cg:Z: tag: `minimap2 -cx <preset> [--secondary=no] contig query`
                       of the stage-3 annotation (TELR_te.py:68-78, 119-132), row f4

Record layout follows minimap2 format.c (mm_write_sam3 / mm_write_paf3): FLAG 0x10 / 0x100 / 0x800, POS = rs + 1, soft clips on
the primary line and hard clips on supplementary and secondary lines, SEQ reverse-complemented for reverse hits, `*` for
secondary lines, tags NM ms AS nn tp cm s1 s2 de, and SA:Z on the non-secondary lines of a read that has several of them
(minimap2's abbreviated form: rname,pos,strand,clip/M/I|D/clip,mapq,NM;).  Not written: rl:i (stated in DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import gzip
import struct

import numpy as np

from . import gather
from .batch import Batch

_ACGT = np.frombuffer(b"ACGTN", np.uint8)
_COMP = np.frombuffer(b"TGCAN", np.uint8)


def _event_de(a, cig) -> float:
    """1 - gap-compressed identity (format.c mm_event_identity)."""
    op, ln = cig & 0xF, cig >> 4
    gaps = (op == 1) | (op == 2)
    n_gap, n_gapo = int(ln[gaps].sum()), int(gaps.sum())
    den = int(a["blen"]) + int(a["n_ambi"]) - n_gap + n_gapo
    return 1.0 - (int(a["mlen"]) / den if den > 0 else 0.0)


def sam_fields(a, cig, qlen: int):
    """(flag, pos0, mapq, cigar words with clips, clip5, clip3, tp) of one alignment record."""
    flag = int(a["flag"])
    hard = bool(flag & 0x900)                       # supplementary, or secondary without --secondary-seq: hard clips (no -Y)
    rev = bool(a["rev"])
    qs, qe = int(a["qs"]), int(a["qe"])
    c5, c3 = (qlen - qe, qs) if rev else (qs, qlen - qe)
    clip_op = 5 if hard else 4
    words = []
    if c5:
        words.append(c5 << 4 | clip_op)
    words += [int(w) for w in cig]
    if c3:
        words.append(c3 << 4 | clip_op)
    tp = "I" if a["inv"] else ("S" if flag & 0x100 else "P")
    return flag, int(a["rs"]), int(a["mapq"]), np.array(words, np.uint32), c5, c3, tp


def _aux(a, cig, tp) -> bytes:
    out = b"".join(t + b"i" + struct.pack("<i", int(v)) for t, v in
                   ((b"NM", int(a["blen"]) - int(a["mlen"]) + int(a["n_ambi"])), (b"ms", a["dp_max"]), (b"AS", a["dp_score"]), (b"nn", a["n_ambi"])))
    out += b"tpA" + tp.encode() + b"cmi" + struct.pack("<i", int(a["cnt"])) + b"s1i" + struct.pack("<i", int(a["score"]))
    if not int(a["flag"]) & 0x100:
        out += b"s2i" + struct.pack("<i", int(a["subsc"]))
    return out + b"def" + struct.pack("<f", _event_de(a, cig))


def _sa_entry(a, qlen: int, ref_name: str) -> str:
    """One SA:Z element the way format.c mm_write_sam3 abbreviates it: clips as S, one M run, the length difference as I or D."""
    qs, qe, rs, re_, rev = int(a["qs"]), int(a["qe"]), int(a["rs"]), int(a["re"]), bool(a["rev"])
    l_i = l_d = 0
    if qe - qs < re_ - rs:
        l_m, l_d = qe - qs, (re_ - rs) - (qe - qs)
    else:
        l_m, l_i = re_ - rs, (qe - qs) - (re_ - rs)
    c5, c3 = (qlen - qe, qs) if rev else (qs, qlen - qe)
    cg = "".join(f"{n}{op}" for n, op in ((c5, "S"), (l_m, "M"), (l_i, "I"), (l_d, "D"), (c3, "S")) if n)
    return f"{ref_name},{rs + 1},{'-' if rev else '+'},{cg},{int(a['mapq'])},{int(a['blen']) - int(a['mlen']) + int(a['n_ambi'])};"


def strand_records(batch: Batch, result, locus: int, strand: int, read_names, ref_name: str = "ctg1"):
    """SAM records (dicts) of the reads of one locus against one contig strand, in minimap2's output order."""
    rb, re_ = int(batch.locus_read_begin[locus]), int(batch.locus_read_begin[locus + 1])
    al = result.alns
    sel = np.nonzero((al["read"] >= rb) & (al["read"] < re_) & (al["strand"] == strand))[0]
    by_read = {}
    for i in sel:
        by_read.setdefault(int(al["read"][i]), []).append(int(i))
    recs = []
    for r in range(rb, re_):
        qlen = int(batch.read_len[r])
        codes = batch.unpack(int(batch.read_off[r]), qlen)
        fw = _ACGT[codes].tobytes()
        name = read_names[r]
        if r not in by_read:
            recs.append(dict(qname=name, flag=4, tid=-1, pos=-1, mapq=0, cigar=np.zeros(0, np.uint32), seq=fw, aux=b""))
            continue
        rc = None
        chimeric = [i for i in by_read[r] if not int(al["flag"][i]) & 0x100]       # primary + supplementary lines (parent == id)
        for i in by_read[r]:
            a = al[i]
            cig = result.cigar_of(i)
            flag, pos, mapq, words, c5, c3, tp = sam_fields(a, cig, qlen)
            if a["rev"] and rc is None:
                rc = _COMP[codes[::-1]].tobytes()
            s = rc if a["rev"] else fw
            if flag & 0x100:
                seq = b""                                   # secondary: '*'
            elif flag & 0x800:
                seq = s[c5: len(s) - c3]                    # supplementary: hard-clipped
            else:
                seq = s
            aux = _aux(a, cig, tp)
            if len(chimeric) > 1 and not flag & 0x100:
                aux += b"SAZ" + "".join(_sa_entry(al[k], qlen, ref_name) for k in chimeric if k != i).encode() + b"\0"
            recs.append(dict(qname=name, flag=flag, tid=0, pos=pos, mapq=mapq, cigar=words, seq=seq, aux=aux))
    return recs


def write_bam(path: str, ref_name: str, ref_len: int, recs, level: int = 6, header_extra: str | None = None):
    """`samtools view -bS | samtools sort | samtools index` on SAM records: coordinate-sorted BAM + .bai (native writer)."""
    arr = (gather.SamRec * max(len(recs), 1))()
    keep = []
    for k, r in enumerate(recs):
        q = r["qname"].encode() if isinstance(r["qname"], str) else r["qname"]
        cg = np.ascontiguousarray(r["cigar"], np.uint32)
        keep += [q, cg]
        x = arr[k]
        x.qname, x.flag, x.tid, x.pos, x.mapq = q, r["flag"], r["tid"], r["pos"], r["mapq"]
        x.cigar, x.n_cigar = (cg.ctypes.data if len(cg) else None), len(cg)
        x.seq, x.l_seq = (r["seq"] if r["seq"] else None), len(r["seq"])
        x.aux, x.l_aux = (r["aux"] if r["aux"] else None), len(r["aux"])
    names = (C.c_char_p * 1)(ref_name.encode())
    lens = (C.c_int32 * 1)(int(ref_len))
    rc = gather.lib().telr_bam_write_sorted(path.encode(), 1, names, lens, header_extra.encode() if header_extra else None, len(recs), arr, level)
    if rc != 0:
        raise gather.IoError(rc, path)


def write_realign_bams(batch: Batch, result, read_names, prefixes, contig_names=None, level: int = 6):
    """`prefixes[l]` + ".realign.sort.bam" (forward contig) and + ".revcomp.realign.sort.bam" for every locus of the batch."""
    paths = []
    for l in range(batch.n_loci):
        L = int(batch.contig_len[l])
        if L <= 0:
            continue
        cname = contig_names[l] if contig_names else "ctg1"
        for strand, sfx in ((0, ""), (1, ".revcomp")):
            p = prefixes[l] + sfx + ".realign.sort.bam"
            write_bam(p, cname, L, strand_records(batch, result, l, strand, read_names, cname), level,
                      "@PG\tID:telr_b200\tPN:telr_b200\tCL:minimap2 -a -x %s\n" % {0: "map-ont", 1: "map-pb", 2: "map-hifi"}.get(batch.preset, "?"))
            paths.append(p)
    return paths


def read_bam(path: str):
    """Parse a BAM back into (refs, records) — test helper and round-trip check (whole file; use gather.BamFile for queries)."""
    with gzip.open(path, "rb") as fh:
        d = fh.read()
    assert d[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", d, 4)
    text = d[8: 8 + l_text].decode()
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", d, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        (ln,) = struct.unpack_from("<i", d, off)
        nm = d[off + 4: off + 4 + ln - 1].decode()
        (lr,) = struct.unpack_from("<i", d, off + 4 + ln)
        refs.append((nm, lr))
        off += 8 + ln
    recs = []
    while off + 4 <= len(d):
        (bs,) = struct.unpack_from("<i", d, off)
        p = off + 4
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", d, p)
        name = d[p + 32: p + 32 + l_name - 1].decode()
        q = p + 32 + l_name
        cig = np.frombuffer(d, "<u4", n_cig, q).copy()
        q += 4 * n_cig
        sb = np.frombuffer(d, np.uint8, (l_seq + 1) // 2, q)
        seq = "".join("=ACMGRSVTWYHKDBN"[(b >> 4) & 15] + "=ACMGRSVTWYHKDBN"[b & 15] for b in sb)[:l_seq]
        q += (l_seq + 1) // 2 + l_seq
        tags = {}
        end = p + bs
        while q < end:
            tag, ty = d[q: q + 2].decode(), chr(d[q + 2])
            q += 3
            if ty == "A":
                tags[tag] = chr(d[q]); q += 1
            elif ty in "cCsSiIf":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}[ty]
                (tags[tag],) = struct.unpack_from(fmt, d, q)
                q += struct.calcsize(fmt)
            elif ty == "Z":
                e = d.index(b"\0", q)
                tags[tag] = d[q:e].decode(); q = e + 1
            else:
                raise ValueError(ty)
        recs.append(dict(qname=name, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cig, seq=seq, tags=tags))
        off = end
    return text, refs, recs


_OPS = "MIDNSHP=X"


def cigar_string(words) -> str:
    return "".join(f"{int(w) >> 4}{_OPS[int(w) & 15]}" for w in words) or "*"


def paf_lines(batch: Batch, result, query_names, target_names, secondary: bool = True, strand: int = 0):
    """PAF lines with the tags of `minimap2 -c` (tp cm s1 s2 NM ms AS nn de cg) for the alignments to one contig strand
    (format.c mm_write_paf3: qname qlen qs qe strand tname tlen rs re mlen blen mapq).  `secondary=False` = --secondary=no."""
    al = result.alns
    locus_of_read = np.repeat(np.arange(batch.n_loci), np.diff(batch.locus_read_begin))
    out = []
    for i in range(len(al)):
        a = al[i]
        if int(a["strand"]) != strand or (not secondary and int(a["flag"]) & 0x100):
            continue
        r = int(a["read"]); l = int(locus_of_read[r])
        cig = result.cigar_of(i)
        tp = "I" if a["inv"] else ("S" if int(a["flag"]) & 0x100 else "P")
        f = [query_names[r], int(batch.read_len[r]), int(a["qs"]), int(a["qe"]), "-" if a["rev"] else "+", target_names[l], int(batch.contig_len[l]),
             int(a["rs"]), int(a["re"]), int(a["mlen"]), int(a["blen"]), int(a["mapq"]),
             f"NM:i:{int(a['blen']) - int(a['mlen']) + int(a['n_ambi'])}", f"ms:i:{int(a['dp_max'])}", f"AS:i:{int(a['dp_score'])}", f"nn:i:{int(a['n_ambi'])}",
             f"tp:A:{tp}", f"cm:i:{int(a['cnt'])}", f"s1:i:{int(a['score'])}"]
        if not int(a["flag"]) & 0x100:
            f.append(f"s2:i:{int(a['subsc'])}")
        f += [f"de:f:{_event_de(a, cig):.4f}", "cg:Z:" + cigar_string(cig)]
        out.append("\t".join(str(x) for x in f))
    return out


# ---------------------------------------------------------------------------------------------------------------
def _batch_of(preset: str, contigs, queries_per_contig):
    """One locus per contig: `contigs` = [(name, bytes)], `queries_per_contig` = [[(name, bytes)]] -> (Batch, query names)."""
    from .batch import PRESETS, name_hash, pack_sequences
    from . import lib
    seqs, names, hashes, lrb = [], [], [], [0]
    for (_, ctg), qs in zip(contigs, queries_per_contig):
        seqs.append(ctg)
        for qn, qseq in qs:
            seqs.append(qseq); names.append(qn); hashes.append(name_hash(qn))
        lrb.append(len(names))
    seq2, nmask, offs, lens = pack_sequences(seqs, lib.lib())
    is_ctg = np.zeros(len(seqs), bool)
    k = 0
    for j in range(len(contigs)):
        is_ctg[k] = True
        k += 1 + lrb[j + 1] - lrb[j]
    n = len(contigs)
    b = Batch(PRESETS[preset], seq2, nmask, offs[~is_ctg].copy(), lens[~is_ctg].copy(), np.array(hashes, np.uint32), np.array(lrb, np.int32),
              offs[is_ctg].copy(), lens[is_ctg].copy(), np.full(n, -1, np.int32), np.full(n, -1, np.int32))
    return b, names


def align_to_bam(contig_name: str, contig: bytes, reads, preset: str, bam_path: str, bw: int | None = 2000, device: int = 0, ctx=None):
    """Row f1 — the polishing alignment of local assembly, `minimap2 -t T -ax <preset> -r2k <contig> <reads> | samtools sort > bam`
    (TELR_assembly.py:199-212): reads = [(name, bytes)], one contig, sorted BAM + index.  Returns the number of SAM records."""
    from . import lib
    b, names = _batch_of(preset, [(contig_name, contig)], [reads])
    own = ctx is None
    ctx = ctx or lib.Context(device)
    try:
        ctx.set_option("bw", bw or 0)
        r = ctx.run(b, want_aln=True)
    finally:
        ctx.set_option("bw", 0)
        if own:
            ctx.close()
    recs = strand_records(b, r, 0, 0, names, contig_name)
    write_bam(bam_path, contig_name, len(contig), recs, header_extra="@PG\tID:telr_b200\tPN:telr_b200\tCL:minimap2 -ax %s%s\n" % (preset, " -r%d" % bw if bw else ""))
    return len(recs)


def align_to_paf(contigs, queries_per_contig, preset: str, secondary: bool = True, device: int = 0, ctx=None):
    """Row f4 — the stage-3 annotation alignments, `minimap2 -cx <preset> [--secondary=no] <contig> <query>` (TELR_te.py:68-78:
    the VCF insertion sequence of a locus against its contig; :119-132: the TE library against every contig): one batch for all
    contigs, PAF lines with cg:Z: in minimap2's order (per contig, per query, hits by rank)."""
    from . import lib
    b, names = _batch_of(preset, contigs, queries_per_contig)
    own = ctx is None
    ctx = ctx or lib.Context(device)
    try:
        r = ctx.run(b, want_aln=True)
    finally:
        if own:
            ctx.close()
    return paf_lines(b, r, names, [c[0] for c in contigs], secondary=secondary, strand=0)
