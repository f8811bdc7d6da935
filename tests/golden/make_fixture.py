"""Build the config-1 fixture (BASELINE.json configs[0]) from the reference's own smoke-test data
/root/reference/test/{ref_38kb,reads,library}.fasta.  The full TELR pipeline cannot run here (NGMLR, Sniffles,
RepeatMasker, wtdbg2, minimap2, samtools are all absent), so stages 1-3 are replaced by a direct construction:
the jockey-bearing read locates the insertion breakpoint on the 38 kb reference by exact 13-mer matches, the
"polished contig" is ref flank + jockey + ref flank, the locus reads are the 18 PacBio CLR reads, preset map-pb.
Writes tests/golden/config1_batch.npz (packed batch) and config1_oracle.json (oracle stage-4 outputs).
Run in the build container only (needs /root/reference)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from telr_b200.batch import Batch, PRESETS, name_hash, pack_sequences  # noqa: E402
from telr_b200.stage4 import read_fasta  # noqa: E402
from tests import orc  # noqa: E402

T = "/root/reference/test/"
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def kmers(s, k=13):
    return {s[i:i + k]: i for i in range(len(s) - k + 1)}


def main():
    ref = read_fasta(T + "ref_38kb.fasta")[0][1].upper()
    te = read_fasta(T + "library.fasta")[0][1].upper()
    reads = [(n, s.upper()) for n, s in read_fasta(T + "reads.fasta")]
    te_k, te_rc = kmers(te), kmers(te.translate(COMP)[::-1])
    ref_k = kmers(ref)
    best = None
    for n, s in reads:
        for strand, tk in ((0, te_k), (1, te_rc)):
            hits = [i for i in range(len(s) - 13) if s[i:i + 13] in tk]
            if best is None or len(hits) > best[0]:
                best = (len(hits), n, s, strand, hits)
    nh, name, s, strand, hits = best
    lo, hi = min(hits), max(hits) + 13
    # reference positions matched by the read flanks right next to the TE segment
    left = [ref_k[s[i:i + 13]] for i in range(max(0, lo - 600), lo - 13) if s[i:i + 13] in ref_k]
    right = [ref_k[s[i:i + 13]] for i in range(hi, min(len(s) - 13, hi + 600)) if s[i:i + 13] in ref_k]
    left_rc = right_rc = []
    bp = int(np.median(left)) if left else int(np.median(right))
    # refine: breakpoint = largest left-flank reference coordinate + 13 (or smallest right-flank one)
    if left:
        bp = max(left) + 13
    elif right:
        bp = min(right)
    te_seq = te if strand == 0 else te.translate(COMP)[::-1]
    fl = 2000
    contig = ref[bp - fl:bp] + te_seq + ref[bp:bp + fl]
    print("TE-bearing read", name, "13-mer hits", nh, "strand", strand, "breakpoint on ref", bp, "contig length", len(contig))
    seqs = [contig] + [r[1] for r in reads]
    seq2, nmask, offs, lens = pack_sequences(seqs)
    b = Batch(PRESETS["map-pb"], seq2, nmask, offs[1:].copy(), lens[1:].copy(), np.array([name_hash(r[0]) for r in reads], np.uint32),
              np.array([0, len(reads)], np.int32), offs[:1].copy(), lens[:1].copy(), np.array([fl], np.int32), np.array([fl + len(te_seq)], np.int32))
    np.savez_compressed(os.path.join(HERE, "config1_batch.npz"), preset=b.preset, seq2=b.seq2, nmask=b.nmask, read_off=b.read_off, read_len=b.read_len,
                        read_hash=b.read_hash, locus_read_begin=b.locus_read_begin, contig_off=b.contig_off, contig_len=b.contig_len,
                        te_start=b.te_start, te_end=b.te_end)
    r = orc.af_run(b, threads=0)
    al = r.alns
    out = dict(cov2x=r.cov2x.tolist(), af=[None if np.isnan(v) else float(v) for v in r.af], dp_cells=int(r.c.dp_cells), n_aln=int(r.c.n_aln),
               depth_sha1=hashlib.sha1(r.depth.tobytes()).hexdigest(), depth_sum=int(r.depth.sum()),
               aln=[[int(a[f]) for f in ("read", "strand", "rs", "re", "qs", "qe", "rev", "flag", "dp_max", "n_cigar")] for a in al],
               note="oracle outputs (CPU restatement; parity with real minimap2 2.22/samtools 1.9 unpinned)")
    json.dump(out, open(os.path.join(HERE, "config1_oracle.json"), "w"), indent=0)
    print("cov2x", out["cov2x"], "af", out["af"], "n_aln", out["n_aln"], "cells", out["dp_cells"])


if __name__ == "__main__":
    main()
