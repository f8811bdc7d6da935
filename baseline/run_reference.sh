#!/usr/bin/env bash
# Regenerate TRUE golden depth with the reference's own tools (minimap2 2.22 + samtools 1.9, envs/telr.yml) and compare
# with this library's per-base depth.  Usage: baseline/run_reference.sh <telr_out_dir> <contig_dir> <preset: map-ont|map-pb>
# Mirrors realignment() TELR_te.py:495-515 and get_median_cov() TELR_te.py:870-884.  Not runnable in the build image
# (the tools are absent there); kept so the unpinned part of the oracle can be pinned wherever they exist.
set -euo pipefail
out=$1; cdir=$2; preset=${3:-map-ont}
command -v minimap2 >/dev/null && command -v samtools >/dev/null || { echo "minimap2/samtools not found" >&2; exit 3; }
for reads in "$out"/telr_reads/*.reads.fa; do
  locus=$(basename "$reads" .reads.fa)
  for sfx in "" ".revcomp"; do
    contig="$cdir/$locus.cns.ctg1$sfx.fa"; [ -s "$contig" ] || continue
    prefix="$out/telr_reads/$locus$sfx"
    minimap2 -a -x "$preset" -v 0 "$contig" "$reads" > "$prefix.sam"
    samtools view -bS "$prefix.sam" > "$prefix.realign.bam"
    samtools sort -@ 1 -o "$prefix.realign.sort.bam" "$prefix.realign.bam"
    samtools index -@ 1 "$prefix.realign.sort.bam"
    samtools depth -aa "$prefix.realign.sort.bam" | cut -f3 > "$prefix.depth.ref"
  done
done
echo "reference depth written to $out/telr_reads/*.depth.ref; compare with: python -m telr_b200.compare_depth $out $cdir $preset"
