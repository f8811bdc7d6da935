"""Compare this library's per-base depth with `samtools depth -aa` output produced by baseline/run_reference.sh.

usage: python -m telr_b200.compare_depth <telr_out_dir> <contig_dir> <map-ont|map-pb>
Needs a B200 (the depth here comes from the CUDA path) and the *.depth.ref files written by the script."""
import glob
import os
import sys

import numpy as np

from . import lib
from .batch import Batch, PRESETS, name_hash, pack_sequences
from .stage4 import read_fasta


def main(out, cdir, preset):
    bad = tot = 0
    ctx = lib.Context(0)
    for reads in sorted(glob.glob(os.path.join(out, "telr_reads", "*.reads.fa"))):
        locus = os.path.basename(reads)[: -len(".reads.fa")]
        contig = os.path.join(cdir, locus + ".cns.ctg1.fa")
        ref = [os.path.join(out, "telr_reads", locus + s + ".depth.ref") for s in ("", ".revcomp")]
        if not (os.path.isfile(contig) and all(os.path.isfile(r) for r in ref)):
            continue
        rd = read_fasta(reads)
        seqs = [read_fasta(contig)[0][1]] + [s for _, s in rd]
        seq2, nmask, offs, lens = pack_sequences(seqs, lib.lib())
        b = Batch(PRESETS[preset], seq2, nmask, offs[1:].copy(), lens[1:].copy(), np.array([name_hash(n) for n, _ in rd], np.uint32),
                  np.array([0, len(rd)], np.int32), offs[:1].copy(), lens[:1].copy(), np.array([-1], np.int32), np.array([-1], np.int32))
        r = ctx.run(b, want_depth=True)
        L = int(lens[0])
        for s in range(2):
            want = np.loadtxt(ref[s], dtype=np.int64, ndmin=1)
            got = r.depth[s * L:(s + 1) * L]
            tot += 1
            if len(want) != L or not (want == got).all():
                bad += 1
                print("MISMATCH", locus, "strand", s, "first diff", int(np.nonzero(want[:L] != got[:len(want)])[0][0]) if len(want) else -1)
    print(f"{tot - bad}/{tot} contig strands identical to minimap2+samtools")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:4]))
