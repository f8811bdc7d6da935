"""End-to-end stage API at scale: writes the files TELR's stage 4 consumes (vcf_parsed, sorted+indexed BAM, raw reads FASTA,
contig FASTAs, annotation BED) for N synthetic loci of a configuration, then calls telr_b200.stage4.get_af and reports where
the time went (native read gather: BAM window queries / raw-read scan / pack / read files; device path; report writing).

  python profiles/e2e_get_af.py --config ont_30k_30x --loci 30000 [--fastq] [--gz] [--no-read-files] [--backend gpu|oracle]

`--backend oracle` replaces the device call with the CPU oracle (small N only; used to try the script without a GPU).
"""
import argparse
import ctypes as C
import json
import logging
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from telr_b200 import gather, stage4, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ont_30k_30x")
    ap.add_argument("--loci", type=int, default=3000)
    ap.add_argument("--backend", default="gpu")
    ap.add_argument("--fastq", action="store_true")
    ap.add_argument("--no-read-files", action="store_true")
    ap.add_argument("--extra-reads", type=float, default=0.0, help="unrelated reads added to the raw read file, as a multiple of the needed ones")
    ap.add_argument("--dir", default=None)
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    root = args.dir or tempfile.mkdtemp(prefix="telr_e2e_")
    os.makedirs(root, exist_ok=True)
    out, cdir = os.path.join(root, "out"), os.path.join(root, "contigs")
    os.makedirs(out, exist_ok=True); os.makedirs(cdir, exist_ok=True)
    t0 = time.time()
    b = synth.generate(args.config, 0, args.loci)
    t_gen = time.time() - t0
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    n = b.n_loci
    # loci every 20 kb on chr1..: windows never overlap
    per_chr = 5000
    chroms = [f"chr{l // per_chr + 1}" for l in range(n)]
    starts = [20000 * (l % per_chr + 1) for l in range(n)]
    names = [f"{c}_{s}_{s + 1}" for c, s in zip(chroms, starts)]
    t0 = time.time()
    with open(os.path.join(root, "vcf.tsv"), "w") as fv, open(os.path.join(root, "te.bed"), "w") as fb:
        for l in range(n):
            fv.write("\t".join([chroms[l], str(starts[l]), str(starts[l] + 1), "100", "10", "0.5", f"id{l}", "ACGT", "rA,rB", "PASS", "0/1", "5", "5", "0.9"]) + "\n")
            fb.write(f"{names[l]}\t{b.te_start[l]}\t{b.te_end[l]}\tjockey\t.\t+\n")
            with open(os.path.join(cdir, names[l] + ".cns.ctg1.fa"), "wb") as fc:
                fc.write(b">ctg1\n" + acgt[b.unpack(int(b.contig_off[l]), int(b.contig_len[l]))].tobytes() + b"\n")
    # raw reads + BAM records (one primary record per read inside its locus window)
    raw = os.path.join(root, "raw.fq" if args.fastq else "raw.fa")
    rng = np.random.default_rng(1)
    recs = (gather.SamRec * b.n_reads)()
    cig_keep, name_keep = [], []
    n_chr = (n + per_chr - 1) // per_chr
    with open(raw, "wb") as fr:
        for l in range(n):
            for r in range(int(b.locus_read_begin[l]), int(b.locus_read_begin[l + 1])):
                rn = f"L{l:06d}_R{r - int(b.locus_read_begin[l]):04d}".encode()
                seq = acgt[b.unpack(int(b.read_off[r]), int(b.read_len[r]))].tobytes()
                if args.fastq:
                    fr.write(b"@" + rn + b" synthetic\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n")
                else:
                    fr.write(b">" + rn + b" synthetic\n" + seq + b"\n")
                cg = np.array([600 << 4], np.uint32)
                cig_keep.append(cg); name_keep.append(rn)
                rec = recs[r]
                rec.qname, rec.flag, rec.tid, rec.pos, rec.mapq = rn, 0, l // per_chr, starts[l] - 500 + (r % 400), 60
                rec.cigar, rec.n_cigar, rec.seq, rec.l_seq, rec.aux, rec.l_aux = cg.ctypes.data, 1, None, 0, None, 0
            for _ in range(int(args.extra_reads * (b.locus_read_begin[l + 1] - b.locus_read_begin[l]))):
                ln = int(rng.integers(2000, 20000))
                fr.write(b">x%d_%d\n" % (l, _) + acgt[rng.integers(0, 4, ln)].tobytes() + b"\n")
    bam = os.path.join(root, "reads.bam")
    ref_names = (C.c_char_p * n_chr)(*[f"chr{i + 1}".encode() for i in range(n_chr)])
    ref_lens = (C.c_int32 * n_chr)(*[20000 * (per_chr + 2)] * n_chr)
    rc = gather.lib().telr_bam_write_sorted(bam.encode(), n_chr, ref_names, ref_lens, None, b.n_reads, recs, 1)
    assert rc == 0, rc
    t_files = time.time() - t0
    if args.no_read_files:
        os.environ["TELR_B200_WRITE_READS"] = "0"
    if args.backend == "oracle":
        from tests import orc

        def fake(batch, devices=None, **k):
            r = orc.af_run(batch, threads=0, want_depth=False, want_aln=False)
            return r.cov2x, r.af, None
        stage4.run_batch = fake
    timing = {}
    orig_gather = gather.gather

    def timed_gather(*a, **k):
        g = orig_gather(*a, **k)
        timing.update(g.timing)
        return g
    gather.gather = timed_gather
    orig_run = stage4.run_batch

    def timed_run(batch, *a, **k):
        t = time.time()
        r = orig_run(batch, *a, **k)
        timing["device_path_s"] = time.time() - t
        timing["batch_reads"], timing["batch_gbases"] = int(batch.n_reads), float(batch.read_len.astype(np.int64).sum()) / 1e9
        return r
    stage4.run_batch = timed_run
    t0 = time.time()
    te_freq = stage4.get_af(out, "s", bam, raw, os.path.join(root, "te.bed"), cdir, os.path.join(root, "vcf.tsv"), 100, 200, 50, 50, "ont" if b.preset == 0 else "pacbio", 16)
    t_all = time.time() - t0
    ok = sum(1 for v in te_freq.values() if v.get("freq") is not None)
    truth = b.meta["truth_af"]
    err = [abs(min(v["freq"], 1) - float(truth[i])) for i, nm in enumerate(names) if (v := te_freq.get(nm)) and v.get("freq") is not None]
    print(json.dumps({"config": args.config, "loci": n, "reads": int(b.n_reads), "raw_reads_file_gb": os.path.getsize(raw) / 1e9, "bam_mb": os.path.getsize(bam) / 1e6,
                      "generate_s": t_gen, "write_inputs_s": t_files, "get_af_s": t_all, "loci_per_s_through_get_af": n / t_all,
                      "gather": timing, "loci_with_af": ok, "mean_abs_af_error_vs_truth": float(np.mean(err)) if err else None}))
    if not args.keep and not args.dir:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
