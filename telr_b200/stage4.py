"""Drop-in replacement for TELR's stage-4 function ``get_af`` (reference src/telr/TELR_te.py:578-838).

Same signature, same returned dict (9 keys per locus), same side files:
  <vcf_parsed>.new                      (TELR_assembly.py:388-415)
  <out>/telr_reads/<locus>.reads.fa     (TELR_assembly.py:429-456, TELR_te.py:615-617)
  <contig_dir>/<locus>.cns.ctg1.revcomp.fa  (TELR_te.py:624-627)
  <vcf_parsed>.freq, <vcf_parsed>.revcomp.freq   (TELR_te.py:677-755)
The body between them - per locus 2 x `minimap2 -a -x <preset>`, up to 8 x `samtools depth`, medians - runs on
the GPU through the C ABI (include/telr_af.h).  Intermediate *.realign.sort.bam files are not produced.

Install into a TELR checkout with:   import telr.telr, telr_b200.stage4;  telr.telr.get_af = telr_b200.stage4.get_af
"""
from __future__ import annotations

import logging
import math
import os
import statistics
import threading
import time

import numpy as np

from . import gather, lib
from .batch import Batch, PRESETS, name_hash

_COMP = bytes.maketrans(b"ACGTacgtNnUuRYKMSWBDHVrykmswbdhv", b"TGCAtgcaNnAaYRMKSWVHDByrmkswvhdb")


def format_time(t):       # TELR_utility.py:34-41
    from datetime import datetime, timedelta
    d = datetime(1, 1, 1) + timedelta(seconds=t)
    if d.hour == 0 and d.minute == 0:
        return "%d seconds" % (d.second)
    elif d.hour == 0 and d.minute != 0:
        return "%d minutes %d seconds" % (d.minute, d.second)
    return "%d hours %d minutes %d seconds" % (d.hour, d.minute, d.second)


def read_fasta(path):
    """[(id, sequence)] of a FASTA or FASTQ file; id = first whitespace-delimited token (Biopython record.id)."""
    recs = []
    with open(path, "rb") as fh:
        first = fh.read(1)
        fh.seek(0)
        if first == b"@":
            while True:
                h = fh.readline()
                if not h:
                    break
                s = fh.readline().strip()
                fh.readline()
                fh.readline()
                recs.append((h[1:].split()[0].decode(), s))
        else:
            name, chunks = None, []
            for line in fh:
                if line.startswith(b">"):
                    if name is not None:
                        recs.append((name, b"".join(chunks)))
                    name, chunks = (line[1:].split() or [b""])[0].decode(), []
                else:
                    chunks.append(line.strip())
            if name is not None:
                recs.append((name, b"".join(chunks)))
    return recs


def median_str(cov2x: int, n: int) -> str:
    """str(statistics.median(values)) given 2*median and the number of values (int for odd n, float for even n)."""
    if cov2x == -1 or cov2x == -4:      # -4: locus outside the implemented scope (contig with > 65 536 minimizers), reported like a missing window
        return "None"
    if cov2x < 0:
        raise statistics.StatisticsError("no median for empty data")   # the reference dies here too (TELR_te.py:882)
    if n % 2 == 1:
        return str(cov2x // 2)
    return str(cov2x / 2)


def window_sizes(L, s, e, fl, fo, ti, to):
    """Number of depth values samtools returns for the 4 windows of one strand (te5p, te3p, flank5p, flank3p)."""
    def n(S, E):
        return max(0, min(E, L) - max(S - 1, 0))
    if ti and s + to + ti < e:
        a, b = n(s + to, s + to + ti), n(e - ti - to, e - to)
    else:
        a = b = n(s, e)
    return a, b, n(s - fl - fo, s - fo), n(e + fo, e + fl + fo)


def te_flank_ratio(te_cov, flank_cov):      # TELR_te.py:564-575
    if te_cov and flank_cov:
        ratio = te_cov / flank_cov
        return None if ratio > 1.5 else ratio
    return None


def combine_af(taf_5p, taf_3p):             # TELR_te.py:818-835
    if taf_5p and taf_3p:
        freq = (taf_5p + taf_3p) / 2 if abs(taf_5p - taf_3p) <= 0.3 else None
    elif taf_5p:
        freq = taf_5p
    elif taf_3p:
        freq = taf_3p
    else:
        freq = None
    if freq:
        if freq > 1:
            freq = 1
    return round(freq, 3) if freq else None


def run_batch(batch: Batch, devices=None, **kw):
    """Run the GPU path on one or several devices (loci sharded by read bases, host-side gather, no collective)."""
    if devices is None:
        env = os.environ.get("TELR_B200_DEVICES")
        devices = [int(x) for x in env.split(",")] if env else [0]
    if len(devices) == 1:
        ctx = lib.Context(devices[0])
        try:
            r = ctx.run(batch, **kw)
            return r.cov2x, r.af, r
        finally:
            ctx.close()
    # A batch packed locus by locus (the native gather's layout) is cut into contiguous ranges of loci of about equal cost, which
    # are views of the same packed arrays; any other layout goes through the LPT partition and re-packs every shard.
    contiguous = batch.is_packed_by_locus()
    shards = partition_contiguous(locus_costs(batch), len(devices)) if contiguous else partition_loci(batch, len(devices))
    cov = np.zeros((batch.n_loci, 8), np.int32)
    af = np.zeros(batch.n_loci, np.float64)
    errs = []

    def work(dev, loci):
        try:
            if not len(loci):
                return
            ctx = lib.Context(dev)
            try:
                sub = batch.slice(loci[0], loci[-1] + 1) if contiguous else batch.subset(loci)
                r = ctx.run(sub)
                cov[loci] = r.cov2x
                af[loci] = r.af
            finally:
                ctx.close()
        except Exception as ex:     # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=work, args=(d, s)) for d, s in zip(devices, shards)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return cov, af, None


def _run_with_bams(batch: Batch, locus_names, out_dir, raw_reads, slice_bases=200_000_000):
    """The device path with alignment records kept, written as the reference's per-locus BAMs (TELR_te.py:502-515).  Loci are
    processed in slices so that the record buffers stay small; read names are recovered from their X31 hashes."""
    from . import realign
    names_by_hash = {}
    for nm, _ in read_fasta_names(raw_reads):
        names_by_hash.setdefault(name_hash(nm), nm)
    cov = np.zeros((batch.n_loci, 8), np.int32)
    af = np.zeros(batch.n_loci, np.float64)
    csum = np.concatenate([[0], np.cumsum(batch.read_len.astype(np.int64))])
    ctx = lib.Context(0)
    try:
        l0 = 0
        while l0 < batch.n_loci:
            l1 = l0 + 1
            while l1 < batch.n_loci and csum[batch.locus_read_begin[l1 + 1]] - csum[batch.locus_read_begin[l0]] <= slice_bases:
                l1 += 1
            sub = batch.subset(range(l0, l1))
            r = ctx.run(sub, want_aln=True)
            cov[l0:l1], af[l0:l1] = r.cov2x, r.af
            rn = [names_by_hash.get(int(h), "read_%08x" % int(h)) for h in sub.read_hash]
            realign.write_realign_bams(sub, r, rn, [os.path.join(out_dir, locus_names[l]) for l in range(l0, l1)])
            l0 = l1
    finally:
        ctx.close()
    return cov, af


def read_fasta_names(path):
    """(id, None) of every record of a FASTA/FASTQ file (plain or gzip) without keeping sequences."""
    import gzip
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with op(path, "rb") as fh:
        first = fh.read(1)
        fh.seek(0)
        if first == b"@":
            while True:
                h = fh.readline()
                if not h:
                    break
                fh.readline(); fh.readline(); fh.readline()
                yield h[1:].split()[0].decode(), None
        else:
            for line in fh:
                if line.startswith(b">"):
                    yield (line[1:].split() or [b""])[0].decode(), None


def locus_costs(batch: Batch) -> np.ndarray:
    """Cost estimate of every locus for the partitioner: read bases + contig length."""
    lrb = batch.locus_read_begin
    csum = np.concatenate([[0], np.cumsum(batch.read_len.astype(np.int64))])
    return (csum[lrb[1:]] - csum[lrb[:-1]]) + batch.contig_len.astype(np.int64)


def partition_costs(cost, n: int):
    """Longest-processing-time-first assignment of items to n shards; every shard is returned in ascending item order."""
    import heapq
    cost = np.asarray(cost, np.int64)
    order = np.argsort(-cost, kind="stable")
    heap = [(0, k) for k in range(n)]
    shards = [[] for _ in range(n)]
    for l in order:
        load, k = heapq.heappop(heap)
        shards[k].append(int(l))
        heapq.heappush(heap, (load + int(cost[l]), k))
    return [sorted(s) for s in shards]


def partition_contiguous(cost, n: int):
    """n contiguous ranges of items with about equal total cost (cuts at the multiples of total / n of the running sum)."""
    cost = np.asarray(cost, np.int64)
    csum = np.cumsum(cost)
    total = int(csum[-1]) if len(csum) else 0
    cuts = [0] + [int(np.searchsorted(csum, total * (k + 1) / n, side="left")) + 1 for k in range(n - 1)] + [len(cost)]
    cuts = np.minimum.accumulate(np.minimum(np.array(cuts)[::-1], len(cost)))[::-1]
    return [list(range(int(cuts[k]), int(max(cuts[k], cuts[k + 1])))) for k in range(n)]


def partition_loci(batch: Batch, n: int):
    """Longest-processing-time-first assignment of loci to n shards by read bases + contig length."""
    return partition_costs(locus_costs(batch), n)


def get_af(out, sample_name, bam, raw_reads, contig_te_annotation, contig_dir, vcf_parsed, flank_intervel_size, flank_offset,
           te_interval_size, te_offset, presets, thread):
    logging.info("Estimating allele frequency...")
    start_time = time.time()
    presets = "map-ont" if presets == "ont" else "map-pb"         # TELR_te.py:595-598

    # ---- prep_assembly_inputs(read_type="all"): reads in the +-1 kb breakpoint window of every locus ----
    # native gather (telr_io): BAI window queries, one streaming pass over the raw reads, packing straight into the batch
    telr_reads_dir = os.path.join(out, "telr_reads")
    os.makedirs(telr_reads_dir, exist_ok=True)
    window = 1000
    rows, chroms, begs, ends, names_l, contigs = [], [], [], [], [], []
    with open(vcf_parsed) as fh:
        for line in fh:
            entry = line.replace("\n", "").split("\t")
            bp = round((int(entry[1]) + int(entry[2])) / 2)        # banker's rounding, as the reference
            rows.append(entry)
            chroms.append(entry[0]); begs.append(max(bp - window, 0)); ends.append(bp + window)
            names_l.append("_".join(entry[0:3]))
    for contig_name in names_l:
        contig = os.path.join(contig_dir, contig_name + ".cns.ctg1.fa")
        if not os.path.isfile(contig):
            print(contig_name + " no assembly")
            contigs.append(None)
            continue
        recs = read_fasta(contig)
        with open(os.path.join(contig_dir, contig_name + ".cns.ctg1.revcomp.fa"), "w") as fo:
            for rid, seq in recs:
                fo.write(">" + rid + "\n" + seq.translate(_COMP)[::-1].decode() + "\n")
        contigs.append(recs[0][1] if recs and len(recs[0][1]) > 0 else None)
    if not os.path.isfile(bam + ".bai") and not os.path.isfile(os.path.splitext(bam)[0] + ".bai"):
        gather.index_bam(bam)                                      # TELR indexes its BAM in stage 1; tolerate a missing .bai
    write_reads = os.environ.get("TELR_B200_WRITE_READS", "1") != "0"
    t_g = time.time()
    g = gather.gather(bam, raw_reads, chroms, begs, ends, contigs, reads_dir=telr_reads_dir if write_reads else None,
                      locus_names=names_l, threads=int(thread) if thread else 0)
    logging.info("Read gather: %d windows, %d unique reads, %d BGZF blocks, BAM %.2f s, raw reads %.2f s (%d scanned), pack %.2f s, "
                 "read files %.2f s (total %.2f s)", len(rows), g.timing["unique_reads"], g.timing["bgzf_blocks"], g.timing["bam_s"],
                 g.timing["reads_s"], g.timing["reads_scanned"], g.timing["pack_s"], g.timing["write_s"], time.time() - t_g)
    with open(vcf_parsed + ".new", "w") as new:
        for entry, n in zip(rows, g.n_names):
            new.write("\t".join(entry) + "\t" + str(int(n)) + "\n")
    loci = [None if c is None else (nm, c) for nm, c in zip(names_l, contigs)]

    logging.info("Perform local realignment...")
    start_time = time.time()
    # contig annotation: last BED row per contig wins (dict overwrite, TELR_te.py:656-675)
    coords = {}
    with open(contig_te_annotation) as fh:
        for line in fh:
            entry = line.replace("\n", "").split("\t")
            contig = os.path.join(contig_dir, entry[0] + ".cns.ctg1.fa")
            if os.path.isfile(contig) and os.stat(contig).st_size != 0:
                coords[entry[0]] = (int(entry[1]), int(entry[2]))

    # ---- the batch: arrays of the native gather + TE coordinates ----
    live = [int(i) for i in g.live_index]
    lrb = g.locus_read_begin
    te_s = np.array([coords.get(loci[i][0], (-1, -1))[0] for i in live], np.int32)
    te_e = np.array([coords.get(loci[i][0], (-1, -1))[1] for i in live], np.int32)
    batch = Batch(PRESETS[presets], g.seq2, g.nmask, g.read_off, g.read_len, g.read_hash, lrb, g.contig_off, g.contig_len, te_s, te_e,
                  int(flank_intervel_size), int(flank_offset), int(te_interval_size or 0), int(te_offset), meta={"owner": g})      # the arrays are views of g's buffers
    if batch.n_loci and int(np.diff(lrb).max()) > 8000:
        logging.warning("a locus has more than 8000 reads: samtools 1.9 depth would cap its pile-up (-d 8000); this path counts every read")
    empty = [j for j in range(len(live)) if lrb[j + 1] == lrb[j]]
    keep_bam = os.environ.get("TELR_B200_KEEP_BAM", "0") == "1"      # the `-k` intermediates <locus>[.revcomp].realign.sort.bam(.bai)
    if batch.n_loci:
        try:
            if keep_bam:
                cov2x, af = _run_with_bams(batch, [names_l[i] for i in live], telr_reads_dir, raw_reads)
            else:
                cov2x, af, _ = run_batch(batch)
        except Exception as e:     # noqa: BLE001   (TELR_te.py:649-652)
            print(e)
            print("Local realignment failed, exiting...")
            raise SystemExit(1)
    else:
        cov2x, af = np.zeros((0, 8), np.int32), np.zeros(0)
    logging.info("Local realignment finished in " + format_time(time.time() - start_time))
    del empty
    for j in np.nonzero((cov2x == -4).any(axis=1))[0]:
        logging.warning("%s: contig outside the implemented scope of the GPU path, coverage reported as None", loci[live[int(j)]][0])

    # ---- .freq / .revcomp.freq, then the AF block exactly as the reference parses them back ----
    pos_of = {i: j for j, i in enumerate(live)}
    fl, fo, ti, to = int(flank_intervel_size), int(flank_offset), int(te_interval_size or 0), int(te_offset)
    te_freq = {}
    for strand, suffix in ((0, ".freq"), (1, ".revcomp.freq")):
        with open(vcf_parsed + suffix, "w") as fo_:
            for i, entry in enumerate(rows):
                if i not in pos_of:
                    continue
                name = loci[i][0]
                if name not in coords:
                    continue
                j = pos_of[i]
                Lc = int(batch.contig_len[j])
                s, e = coords[name]
                if strand:
                    s, e = Lc - e, Lc - s
                ns = window_sizes(Lc, s, e, fl, fo, ti, to)
                vals = [median_str(int(cov2x[j, strand * 4 + k]), ns[k]) for k in range(4)]
                fo_.write("\t".join(entry + vals) + "\n")
                cov = [None if v == "None" else float(v) for v in vals]
                d = te_freq.setdefault(name, {})
                sfx = "_rc" if strand else ""
                d["te_5p_cov" + sfx], d["te_3p_cov" + sfx], d["flank_5p_cov" + sfx], d["flank_3p_cov" + sfx] = cov
                if strand:
                    freq = combine_af(te_flank_ratio(d["te_5p_cov"], d["flank_5p_cov"]), te_flank_ratio(d["te_5p_cov_rc"], d["flank_5p_cov_rc"]))
                    # the returned value is the reference's own formula on the coverage integers; the device value (fp64, before
                    # clamp/round) is only cross-checked against it, to the tolerance the path promises (1e-9 relative)
                    g = af[j]
                    t5, t3 = te_flank_ratio(d["te_5p_cov"], d["flank_5p_cov"]), te_flank_ratio(d["te_5p_cov_rc"], d["flank_5p_cov_rc"])
                    raw = ((t5 + t3) / 2 if abs(t5 - t3) <= 0.3 else None) if (t5 and t3) else (t5 or t3 or None)
                    ok = math.isnan(g) if raw is None else (not math.isnan(g) and abs(g - raw) <= 1e-9 * abs(raw))
                    if not ok:
                        raise RuntimeError(f"{name}: device AF {g!r} disagrees with the reference formula {raw!r}")
                    d["freq"] = freq
    logging.info("Allele frequency estimation finished in " + format_time(time.time() - start_time))
    return te_freq
