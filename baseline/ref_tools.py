"""The reference's own CPU path for stage 4, run with the real third-party binaries when they exist.

Used ONLY by `bench.py --impl reference` / `cpu_baseline` when `minimap2` and `samtools` are found on PATH or under
baseline/_ref/bin (they are not part of this image).  It re-enacts, command for command, what
bergmanlab/TELR does for every locus of a batch:

  realignment()      TELR_te.py:495-515   minimap2 -a -x <preset> -v 0 contig reads | samtools view -bS | sort | index
                     run through a multiprocessing Pool of `thread` workers, forward contigs first, then reverse
                     complements (TELR_te.py:640-654)
  get_median_cov()   TELR_te.py:870-884   samtools depth -aa -r chr:S-E  + statistics.median
  get_te_cov / get_flank_cov window rules TELR_te.py:841-867, 518-550 (serial loops, TELR_te.py:677-755)
  AF block           TELR_te.py:785-835

TELR itself cannot be imported here (Biopython / pysam are absent), so the few lines of Python glue are restated; the
heavy arithmetic is the unmodified binaries.  Returns {locus index: dict of the 8 coverages + freq}.
"""
from __future__ import annotations

import os
import shutil
import statistics
import subprocess
import tempfile
from multiprocessing import Pool

import numpy as np

_NT = np.frombuffer(b"ACGTN", np.uint8)
_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _realign(args):        # TELR_te.py:495-515
    mm2, sam_tools, preset, contig, reads, prefix = args
    sam = prefix + ".realign.sam"
    with open(sam, "w") as out:
        subprocess.call([mm2, "-a", "-x", preset, "-v", "0", contig, reads], stdout=out)
    bam = prefix + ".realign.bam"
    with open(bam, "w") as out:
        subprocess.call([sam_tools, "view", "-bS", sam], stdout=out)
    sorted_bam = prefix + ".realign.sort.bam"
    subprocess.call([sam_tools, "sort", "-o", sorted_bam, bam])
    subprocess.call([sam_tools, "index", sorted_bam])
    os.remove(sam)
    os.remove(bam)
    return sorted_bam


def _median_cov(sam_tools, bam, chrom, start, end):     # TELR_te.py:870-884
    out = subprocess.check_output([sam_tools, "depth", "-aa", "-r", f"{chrom}:{start}-{end}", bam]).decode()
    cov = [int(line.split("\t")[2]) for line in out.splitlines() if line]
    return statistics.median(cov)


def _strand_cov(sam_tools, bam, chrom, L, s, e, fl, fo, ti, to):
    if ti and s + to + ti < e:          # get_te_cov, TELR_te.py:841-867
        te5 = _median_cov(sam_tools, bam, chrom, s + to, s + to + ti)
        te3 = _median_cov(sam_tools, bam, chrom, e - ti - to, e - to)
    else:
        te5 = te3 = _median_cov(sam_tools, bam, chrom, s, e)
    fl5 = _median_cov(sam_tools, bam, chrom, s - fl - fo, s - fo) if s - fl - fo >= 0 else None      # get_flank_cov, TELR_te.py:518-550
    fl3 = _median_cov(sam_tools, bam, chrom, e + fo, e + fl + fo) if e + fl + fo <= L else None
    return te5, te3, fl5, fl3


def run(batch, minimap2: str, samtools: str, threads: int = 1, keep: bool = False):
    from telr_b200.stage4 import combine_af, te_flank_ratio
    preset = {0: "map-ont", 1: "map-pb", 2: "map-hifi"}[batch.preset]
    tmp = tempfile.mkdtemp(prefix="telr_ref_")
    jobs_fw, jobs_rc, meta = [], [], []
    for l in range(batch.n_loci):
        L = int(batch.contig_len[l])
        if L <= 0:
            meta.append(None)
            continue
        ctg = _NT[batch.unpack(int(batch.contig_off[l]), L)].tobytes()
        name = f"locus{l}"
        fw, rc, rd = os.path.join(tmp, name + ".fa"), os.path.join(tmp, name + ".revcomp.fa"), os.path.join(tmp, name + ".reads.fa")
        with open(fw, "wb") as fh:
            fh.write(b">" + name.encode() + b"\n" + ctg + b"\n")
        with open(rc, "wb") as fh:
            fh.write(b">" + name.encode() + b"\n" + ctg.translate(_COMP)[::-1] + b"\n")
        with open(rd, "wb") as fh:
            for r in range(int(batch.locus_read_begin[l]), int(batch.locus_read_begin[l + 1])):
                fh.write(b">r%d\n" % r + _NT[batch.unpack(int(batch.read_off[r]), int(batch.read_len[r]))].tobytes() + b"\n")
        jobs_fw.append((minimap2, samtools, preset, fw, rd, os.path.join(tmp, name)))
        jobs_rc.append((minimap2, samtools, preset, rc, rd, os.path.join(tmp, name + ".revcomp")))
        meta.append((name, L))
    with Pool(processes=max(1, threads)) as pool:       # TELR_te.py:643-648 (forward), then the reverse complements
        bams_fw = pool.map(_realign, jobs_fw)
        bams_rc = pool.map(_realign, jobs_rc)
    out, k = {}, 0
    fl, fo, ti, to = batch.flank_len, batch.flank_off, batch.te_len, batch.te_off
    for l, m in enumerate(meta):
        if m is None:
            continue
        name, L = m
        s, e = int(batch.te_start[l]), int(batch.te_end[l])
        if s >= 0:
            c = _strand_cov(samtools, bams_fw[k], name, L, s, e, fl, fo, ti, to)
            c_rc = _strand_cov(samtools, bams_rc[k], name, L, L - e, L - s, fl, fo, ti, to)
            freq = combine_af(te_flank_ratio(c[0], c[2]), te_flank_ratio(c_rc[0], c_rc[2]))
            out[l] = {"fw": c, "rc": c_rc, "freq": freq}
        k += 1
    if not keep:
        shutil.rmtree(tmp, ignore_errors=True)
    return out
