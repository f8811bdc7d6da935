"""Host-side pieces without a GPU: C-ABI library loads and exports every symbol of include/telr_af.h, fails
loudly without a device, BAM I/O, batch packing, the get_af drop-in's file handling, and the N>1 sharding path
(world_size-2 gloo)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from telr_b200 import gather, lib, realign, stage4, synth
from tests import bamio
from telr_b200.batch import Batch, PRESETS, name_hash, pack_sequences
from tests import orc, util


def test_cabi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(util.ROOT, "include", "telr_af.h")).read()
    declared = set(re.findall(r"\b(telr_[a-z_0-9]+)\s*\(", hdr))
    L = C.CDLL(lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(lib.EXPORTS)
    assert lib.lib().telr_af_version() >= 100
    assert lib.lib().telr_af_strerror(-4).decode().startswith("no sm_100")


def test_io_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(util.ROOT, "include", "telr_io.h")).read()
    declared = set(re.findall(r"\b(telr_[a-z_0-9]+)\s*\(", hdr))
    L = C.CDLL(gather.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(gather.EXPORTS)
    assert gather.lib().telr_io_strerror(-5).decode().startswith("BAM index")


def test_chunk_planner(built):
    """How telr_af_run cuts a batch into chunks of loci (host logic, no device): whole loci, equal shares, edge cases."""
    rng = np.random.default_rng(5)
    n_loci = 200
    per = rng.integers(0, 80, n_loci)
    per[[3, 50, 199]] = 0                                           # loci without reads, also at the very end
    lrb = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    read_len = rng.integers(1000, 30000, int(lrb[-1])).astype(np.int32)
    locus_bases = np.add.reduceat(np.concatenate([read_len, [0]]).astype(np.int64), lrb[:-1]) * (per > 0)
    total = int(read_len.astype(np.int64).sum())
    assert int(locus_bases.sum()) == total
    for budget in (1, 10_000, 3_000_000, total // 3, total // 2 + 1, total, 10 * total):
        cuts = lib.plan_chunks(read_len, lrb, budget)
        assert cuts[0] == 0 and cuts[-1] == n_loci and (np.diff(cuts) >= 1).all()
        n_want = max(1, -(-total // budget))
        sizes = np.array([int(locus_bases[a:b].sum()) for a, b in zip(cuts[:-1], cuts[1:])])
        if budget >= int(locus_bases.max()):
            assert len(sizes) <= n_want + 1
            assert sizes.max() <= total / n_want + locus_bases.max()     # a chunk overshoots its share by less than one locus
        if budget >= total:
            assert len(sizes) == 1
    # a single locus above the budget is one chunk; an empty batch has no chunks
    assert lib.plan_chunks(np.array([5000, 7000], np.int32), np.array([0, 2], np.int32), 100).tolist() == [0, 1]
    assert lib.plan_chunks(np.zeros(0, np.int32), np.array([0], np.int32), 100).tolist() == [0]
    assert lib.plan_chunks(np.zeros(0, np.int32), np.array([0, 0, 0], np.int32), 100).tolist() == [0, 2]


def test_no_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.TelrError) as ei:
        lib.Context(0)
    assert ei.value.code == -4          # TELR_ENODEV: there is no CPU fallback


def test_pack_and_hash_helpers(built):
    L = lib.lib()
    seqs = [b"ACGTNNacgtRYK" * 7, b"T" * 64, b"G"]
    s2a, nma, offa, lena = pack_sequences(seqs)             # numpy path
    s2b, nmb, offb, lenb = pack_sequences(seqs, L)          # C helper
    assert (s2a == s2b).all() and (nma == nmb).all() and (offa == offb).all()
    b = Batch(0, s2a, nma, offa[1:], lena[1:], np.zeros(2, np.uint32), np.array([0, 2], np.int32), offa[:1], lena[:1],
              np.array([1], np.int32), np.array([2], np.int32))
    u = b.unpack(0, len(seqs[0]))
    assert u[:6].tolist() == [0, 1, 2, 3, 4, 4] and u[10] == 4
    for nm in ("L000001_R0003", "m150806_161050_42131/92222/1146_17405", "x"):
        assert L.telr_name_hash(nm.encode()) == name_hash(nm) == orc.lib().orc_name_hash(nm.encode())


def test_synth_is_deterministic_and_shardable(built):
    a = synth.generate("ont_3k_50x", 0, 6, depth=8)
    b1 = synth.generate("ont_3k_50x", 0, 3, depth=8)
    b2 = synth.generate("ont_3k_50x", 3, 3, depth=8)
    assert (a.contig_len == np.concatenate([b1.contig_len, b2.contig_len])).all()
    assert (a.read_hash == np.concatenate([b1.read_hash, b2.read_hash])).all()
    sub = a.subset([3, 4, 5])
    assert (sub.seq2 == b2.seq2).all() and (sub.read_len == b2.read_len).all() and (sub.te_start == b2.te_start).all()


def test_bam_roundtrip_and_fetch(tmp_path):
    recs = [dict(name="r1", ref_id=0, pos=100, cigar=[(0, 50), (2, 10), (0, 40)], seq_len=90),
            dict(name="r2", ref_id=0, pos=900, cigar=[(4, 5), (0, 300)], seq_len=305, flag=256),
            dict(name="r3", ref_id=0, pos=1200, cigar=[(0, 10)], seq_len=10, flag=2048),
            dict(name="r4", ref_id=1, pos=0, cigar=[(0, 10)], seq_len=10)]
    p = str(tmp_path / "t.bam")
    bamio.write_bam(p, [("chrA", 5000), ("chrB", 100)], recs)
    idx = bamio.BamIndex(p)
    assert idx.refs == [("chrA", 5000), ("chrB", 100)]
    assert list(idx.fetch("chrA", 0, 100)) == []                 # r1 starts at 100: half-open window
    assert list(idx.fetch("chrA", 199, 200)) == ["r1"]           # reference span 100 (50M10D40M) -> [100, 200)
    assert list(idx.fetch("chrA", 200, 901)) == ["r2"]           # secondary records count too
    assert sorted(idx.fetch("chrA", 0, 5000)) == ["r1", "r2", "r3"]
    with pytest.raises(ValueError):
        list(idx.fetch("chrZ", 0, 1))


def _random_bam(path, rng, n_rec, refs):
    """A coordinate-sorted BAM whose records fall into every BAI bin level (spans from 1 base to several Mbase)."""
    recs = []
    for i in range(n_rec):
        tid = int(rng.integers(0, len(refs)))
        L = refs[tid][1]
        span = int(rng.choice([1, 30, 900, 20000, 300000, 3000000]))
        span = max(1, min(span, L - 1))
        pos = int(rng.integers(0, L - span))
        kind = rng.random()
        if kind < 0.1:
            cigar = []                                            # placed but without a CIGAR: counts as length 1
        elif kind < 0.5:
            cigar = [(4, 7), (0, span)]
        else:
            a = max(1, span // 3)
            cigar = [(0, a), (2, span - a) if span - a > 0 else (1, 3), (1, 5)] if span > 1 else [(0, 1)]
        recs.append(dict(name=f"q{i}_{'x' * int(rng.integers(0, 40))}", ref_id=tid, pos=pos, cigar=cigar, seq_len=0,
                         flag=int(rng.choice([0, 16, 256, 2048, 2064]))))
    recs.sort(key=lambda r: (r["ref_id"], r["pos"]))
    bamio.write_bam(path, refs, recs)
    return recs


def test_native_bam_fetch_matches_the_pure_python_reader(built, tmp_path):
    """telr_bam_fetch (BGZF random access through the .bai written by telr_bam_index_build) against bamio.BamIndex, which
    inflates the whole file and scans it: same names in the same order for random windows; and an indexed query touches
    a handful of BGZF blocks, not the file."""
    rng = np.random.default_rng(42)
    refs = [("chr2L", 23_000_000), ("chrX", 400_000_000), ("tiny", 40)]
    p = str(tmp_path / "r.bam")
    _random_bam(p, rng, 6000, refs)
    gather.index_bam(p)
    ref = bamio.BamIndex(p)
    nat = gather.BamFile(p)
    assert nat.refs == refs
    total_blocks = os.path.getsize(p) // 20000
    for _ in range(300):
        chrom, L = refs[int(rng.integers(0, 3))]
        w = int(rng.choice([1, 2000, 2000, 100000, 5_000_000]))
        s = int(rng.integers(0, max(1, L - 1)))
        want = list(ref.fetch(chrom, s, s + w))
        b0 = nat.blocks_inflated
        got = nat.fetch(chrom, s, s + w)
        assert got == want, (chrom, s, w, len(got), len(want))
        if w <= 2000:
            assert nat.blocks_inflated - b0 <= 12
    assert nat.fetch("chrX", 0, 400_000_000) == list(ref.fetch("chrX", 0, 400_000_000))
    assert nat.fetch("tiny", 50, 60) == [] and nat.fetch("chr2L", 10, 10) == []
    with pytest.raises(ValueError):
        nat.fetch("chrZ", 0, 1)
    nat.close()
    os.remove(p + ".bai")
    with pytest.raises(gather.IoError):
        gather.BamFile(p)


@pytest.mark.parametrize("fmt", ["fasta", "fastq", "fasta.gz"])
def test_native_gather_packs_the_same_batch_as_the_python_path(built, tmp_path, fmt):
    """telr_gather_run against the previous pure-Python gather (BamIndex.fetch + dict of reads + pack_sequences)."""
    import gzip
    rng = np.random.default_rng(7)
    n_reads, n_loci = 300, 12
    alpha = np.frombuffer(b"ACGTNacgtnRY", np.uint8)
    reads = {f"read{i}/ccs": bytes(alpha[rng.choice(len(alpha), int(rng.integers(1, 3000)), p=[.23, .23, .23, .23, .01, .01, .01, .01, .01, .01, .01, .01])]) for i in range(n_reads)}
    names = list(reads)
    recs = []
    for i, nm in enumerate(names):
        for _ in range(int(rng.integers(1, 3))):                  # supplementary / secondary records repeat the name
            recs.append(dict(name=nm, ref_id=0, pos=int(rng.integers(0, 60000)), cigar=[(0, int(rng.integers(50, 4000)))], seq_len=0, flag=int(rng.choice([0, 256, 2048]))))
    recs.sort(key=lambda r: r["pos"])
    bam = str(tmp_path / "g.bam")
    bamio.write_bam(bam, [("chr1", 100000)], recs)
    gather.index_bam(bam)
    order = rng.permutation(n_reads)
    raw = str(tmp_path / ("raw." + fmt))
    if fmt == "fastq":
        body = b"".join(b"@" + names[i].encode() + b" desc\n" + reads[names[i]] + b"\n+\n" + b"I" * len(reads[names[i]]) + b"\n" for i in order)
    else:
        body = b"".join(b">" + names[i].encode() + b" desc\n" + b"\n".join(reads[names[i]][k:k + 70] for k in range(0, len(reads[names[i]]), 70)) + b"\n" for i in order)
    with (gzip.open(raw, "wb") if fmt.endswith(".gz") else open(raw, "wb")) as fh:
        fh.write(body)
    begs = [int(x) for x in rng.integers(0, 58000, n_loci)]
    ends = [b + 2000 for b in begs]
    contigs = [bytes(alpha[rng.integers(0, 5, int(rng.integers(100, 900)))]) if l % 4 != 3 else None for l in range(n_loci)]
    rdir = tmp_path / "reads"; rdir.mkdir()
    g = gather.gather(bam, raw, ["chr1"] * n_loci, begs, ends, contigs, reads_dir=str(rdir), locus_names=[f"L{l}" for l in range(n_loci)], threads=3)
    idx = bamio.BamIndex(bam)
    per_locus = [sorted(set(idx.fetch("chr1", b, e))) for b, e in zip(begs, ends)]
    assert g.n_names.tolist() == [len(x) for x in per_locus]
    live = [l for l in range(n_loci) if contigs[l]]
    assert g.live_index.tolist() == live
    seqs, hashes, lrb = [], [], [0]
    for l in live:
        seqs.append(contigs[l])
        for nm in per_locus[l]:
            seqs.append(reads[nm]); hashes.append(name_hash(nm))
        lrb.append(len(hashes))
    seq2, nmask, offs, lens = pack_sequences(seqs)
    assert g.n_bases == len(seq2) * 16 and (g.seq2 == seq2).all() and (g.nmask == nmask).all()
    is_ctg = np.zeros(len(seqs), bool); k = 0
    for j in range(len(live)):
        is_ctg[k] = True; k += 1 + lrb[j + 1] - lrb[j]
    assert (g.read_off == offs[~is_ctg]).all() and (g.read_len == lens[~is_ctg]).all() and (g.read_hash == np.array(hashes, np.uint32)).all()
    assert (g.contig_off == offs[is_ctg]).all() and (g.contig_len == lens[is_ctg]).all() and (g.locus_read_begin == np.array(lrb)).all()
    for l in range(n_loci):                                        # read files exist for every locus, contig or not
        fa = (rdir / f"L{l}.reads.fa").read_bytes()
        assert fa == b"".join(b">" + nm.encode() + b"\n" + reads[nm] + b"\n" for nm in per_locus[l])
    assert g.timing["unique_reads"] == len(set(sum(per_locus, [])))
    g.free()
    # a read named in the BAM but absent from the raw reads: KeyError like SeqIO.index(...).get_raw; unknown contig: ValueError like pysam
    open(raw, "wb").write(b">someone_else\nACGT\n")
    with pytest.raises(KeyError):
        gather.gather(bam, raw, ["chr1"], [0], [60000], [b"ACGT"])
    with pytest.raises(ValueError):
        gather.gather(bam, raw, ["chrNope"], [0], [10], [b"ACGT"])


def _make_stage3_artifacts(tmp_path, b: Batch, drop_contig=None, drop_annot=None):
    """Write the files get_af() consumes (vcf_parsed, BAM, raw reads, contigs, BED) for a synthetic batch."""
    out = tmp_path / "out"; cdir = tmp_path / "contigs"; out.mkdir(); cdir.mkdir()
    acgt = np.array(list(b"ACGTN"), np.uint8)
    rows, bed, recs, fa = [], [], [], []
    for l in range(b.n_loci):
        start = 10000 * (l + 1)
        name = f"chr1_{start}_{start + 1}"
        rows.append("\t".join(["chr1", str(start), str(start + 1), "100", "10", "0.5", f"id{l}", "ACGT", "rA,rB", "PASS", "0/1", "5", "5", "0.9"]))
        if l != drop_contig:
            (cdir / f"{name}.cns.ctg1.fa").write_bytes(b">ctg1\n" + acgt[b.unpack(int(b.contig_off[l]), int(b.contig_len[l]))].tobytes() + b"\n")
        if l != drop_annot:
            bed.append(f"{name}\t1\t2\tdummy\t.\t+")            # overwritten by the next row: last row per contig wins
            bed.append(f"{name}\t{b.te_start[l]}\t{b.te_end[l]}\tjockey\t.\t+")
        for r in range(b.locus_read_begin[l], b.locus_read_begin[l + 1]):
            rn = f"L{l:06d}_R{r - b.locus_read_begin[l]:04d}"
            recs.append(dict(name=rn, ref_id=0, pos=start - 500 + (r % 7), cigar=[(0, 600)], seq_len=0))
            fa.append(b">" + rn.encode() + b" extra words\n" + acgt[b.unpack(int(b.read_off[r]), int(b.read_len[r]))].tobytes() + b"\n")
    recs.append(dict(name="far_away", ref_id=0, pos=5, cigar=[(0, 50)], seq_len=0))
    fa.append(b">far_away\nACGT\n")
    recs.sort(key=lambda r: r["pos"])
    bamio.write_bam(str(tmp_path / "reads.bam"), [("chr1", 10 ** 6)], recs)
    (tmp_path / "raw.fa").write_bytes(b"".join(fa))
    (tmp_path / "vcf.tsv").write_text("\n".join(rows) + "\n")
    (tmp_path / "te.bed").write_text("\n".join(bed) + "\n")
    return dict(out=str(out), sample_name="s", bam=str(tmp_path / "reads.bam"), raw_reads=str(tmp_path / "raw.fa"),
                contig_te_annotation=str(tmp_path / "te.bed"), contig_dir=str(cdir), vcf_parsed=str(tmp_path / "vcf.tsv"),
                flank_intervel_size=100, flank_offset=200, te_interval_size=50, te_offset=50, presets="ont", thread=1)


def test_get_af_host_logic_with_oracle_backend(built, tmp_path, monkeypatch):
    """The drop-in's file handling and dict/.freq construction, with the device call replaced by the oracle
    (the product path itself is exercised on the GPU in test_gpu.py::test_get_af_dropin)."""
    b = synth.generate("ont_3k_50x", 0, 4, depth=10)
    kw = _make_stage3_artifacts(tmp_path, b, drop_contig=1, drop_annot=2)
    seen = {}

    def fake_run_batch(batch, devices=None, **k):
        seen["batch"] = batch
        r = orc.af_run(batch, threads=0, want_depth=False, want_aln=False)
        return r.cov2x, r.af, None
    monkeypatch.setattr(stage4, "run_batch", fake_run_batch)
    te_freq = stage4.get_af(**kw)
    names = [f"chr1_{10000 * (l + 1)}_{10000 * (l + 1) + 1}" for l in range(4)]
    assert set(te_freq) == {names[0], names[3]}                  # locus 1 has no contig, locus 2 no annotation
    gb = seen["batch"]
    assert gb.n_loci == 3 and gb.preset == PRESETS["map-ont"]
    keep = [0, 2, 3]
    ref = orc.af_run(b.subset(keep), threads=0, want_depth=False, want_aln=False)
    # reads were re-gathered through BAM + FASTA in sorted-name order = generator order: identical batch content
    assert (gb.read_len == b.subset(keep).read_len).all() and (gb.read_hash == b.subset(keep).read_hash).all()
    for j, l in ((0, 0), (2, 3)):
        d = te_freq[names[l]]
        assert set(d) == {"te_5p_cov", "te_3p_cov", "flank_5p_cov", "flank_3p_cov", "te_5p_cov_rc", "te_3p_cov_rc", "flank_5p_cov_rc", "flank_3p_cov_rc", "freq"}
        for k, key in enumerate(["te_5p_cov", "te_3p_cov", "flank_5p_cov", "flank_3p_cov", "te_5p_cov_rc", "te_3p_cov_rc", "flank_5p_cov_rc", "flank_3p_cov_rc"]):
            c2 = int(ref.cov2x[j, k])
            assert d[key] == (None if c2 == -1 else c2 / 2)
        g = ref.af[j]
        assert d["freq"] == (None if np.isnan(g) else round(1 if g > 1 else g, 3))
    # side files
    new = open(kw["vcf_parsed"] + ".new").read().splitlines()
    assert len(new) == 4 and all(len(x.split("\t")) == 15 for x in new)
    assert [int(x.split("\t")[14]) for x in new] == np.diff(b.locus_read_begin).tolist()
    freq = open(kw["vcf_parsed"] + ".freq").read().splitlines()
    rc = open(kw["vcf_parsed"] + ".revcomp.freq").read().splitlines()
    assert len(freq) == 2 and len(rc) == 2 and all(len(x.split("\t")) == 18 for x in freq + rc)
    assert os.path.isfile(os.path.join(kw["out"], "telr_reads", names[1] + ".reads.fa"))     # written even without a contig
    rcfa = open(os.path.join(kw["contig_dir"], names[0] + ".cns.ctg1.revcomp.fa")).read().splitlines()
    fw = open(os.path.join(kw["contig_dir"], names[0] + ".cns.ctg1.fa")).read().splitlines()
    assert rcfa[0] == ">ctg1" and rcfa[1] == fw[1][::-1].translate(str.maketrans("ACGT", "TGCA"))


def test_partition_is_balanced_and_complete(built):
    b = synth.generate("ont_3k_50x", 0, 16, depth=6)
    for n in (1, 2, 3, 8):
        sh = stage4.partition_loci(b, n)
        assert sorted(sum(sh, [])) == list(range(16))
        cost = [sum(int(b.read_len[b.locus_read_begin[l]:b.locus_read_begin[l + 1]].sum()) for l in s) for s in sh]
        assert max(cost) <= 1.6 * (sum(cost) / n) + max(int(b.read_len.sum()) // 16, 1)


def test_contiguous_shards_are_views_of_the_batch(built):
    """run_batch over several devices cuts a batch that is packed locus by locus into contiguous ranges of about equal cost;
    a range is a view of the packed arrays (Batch.slice) and holds exactly what the re-packing subset() holds."""
    b = synth.generate("ont_3k_50x", 0, 17, depth=6)
    assert b.is_packed_by_locus()
    cost = stage4.locus_costs(b)
    for n in (1, 2, 3, 8, 17, 20):
        sh = stage4.partition_contiguous(cost, n)
        assert len(sh) == n and sum(sh, []) == list(range(17))
        loads = [int(cost[s].sum()) for s in sh if s]
        if n <= 8:
            assert max(loads) <= cost.sum() / n + cost.max()
        for s_ in sh:
            if not s_:
                continue
            v, w = b.slice(s_[0], s_[-1] + 1), b.subset(s_)
            v.validate()
            assert np.shares_memory(v.seq2, b.seq2)
            assert (v.read_len == w.read_len).all() and (v.read_hash == w.read_hash).all() and (v.locus_read_begin == w.locus_read_begin).all()
            assert (v.contig_len == w.contig_len).all() and (v.te_start == w.te_start).all()
            for r in range(0, v.n_reads, 7):
                assert (v.unpack(int(v.read_off[r]), int(v.read_len[r])) == w.unpack(int(w.read_off[r]), int(w.read_len[r]))).all()
            for l in range(v.n_loci):
                assert (v.unpack(int(v.contig_off[l]), int(v.contig_len[l])) == w.unpack(int(w.contig_off[l]), int(w.contig_len[l]))).all()
            ro_v = orc.af_run(v, threads=0, want_depth=False, want_aln=False) if n == 3 else None
            if ro_v is not None:
                ro_w = orc.af_run(w, threads=0, want_depth=False, want_aln=False)
                assert (ro_v.cov2x == ro_w.cov2x).all()
    shuffled = b.subset([3, 1, 2])
    assert shuffled.is_packed_by_locus()          # re-packed in the requested order
    b2 = Batch(b.preset, b.seq2, b.nmask, b.read_off, b.read_len, b.read_hash, b.locus_read_begin, b.contig_off[::-1].copy(), b.contig_len[::-1].copy(),
               b.te_start, b.te_end)
    assert not b2.is_packed_by_locus()


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synth.generate("ont_3k_50x", 0, 6, depth=6)
    mine = stage4.partition_loci(b, world)[rank]
    r = orc.af_run(b.subset(mine), threads=1, want_depth=False, want_aln=False)      # stand-in for the device call
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, r.cov2x.tolist()))
    if rank == 0:
        cov = np.zeros((b.n_loci, 8), np.int32)
        for loci, c in gathered:
            cov[loci] = np.array(c, np.int32)
        q.put(cov.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(built):
    """N>1 path: loci sharded by rank, independent work, host-side gather in locus order (no data-path collective)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    cov = np.array(q.get(timeout=300), np.int32)
    [p.join(60) for p in ps]
    b = synth.generate("ont_3k_50x", 0, 6, depth=6)
    ref = orc.af_run(b, threads=0, want_depth=False, want_aln=False)
    assert (cov == ref.cov2x).all()


def _depth_from_records(recs, L):
    """`samtools depth -aa` over BAM records: every record that is not SECONDARY or UNMAPPED, M/=/X bases only."""
    d = np.zeros(L + 1, np.int64)
    for r in recs:
        if r["flag"] & 0x104:
            continue
        pos = r["pos"]
        for w in r["cigar"].tolist():
            op, ln = w & 15, w >> 4
            if op in (0, 7, 8):
                d[pos] += 1; d[pos + ln] -= 1; pos += ln
            elif op in (2, 3):
                pos += ln
    return np.cumsum(d)[:L]


def test_realign_bam_records_reproduce_the_depth(built, tmp_path):
    """Row f2 on the host: SAM records built from alignment records (here the oracle's), written as sorted BAM + BAI by the
    native writer, read back: coordinate order, FLAG/CIGAR/SEQ consistency of `minimap2 -a` output, and `samtools depth`
    recomputed from the BAM equals the depth the path reports for that contig strand."""
    b = synth.generate("ont_3k_50x", 7, 3, depth=12)
    ro = orc.af_run(b, threads=0)
    names = [f"read{r}" for r in range(b.n_reads)]
    paths = realign.write_realign_bams(b, ro, names, [str(tmp_path / f"L{l}") for l in range(b.n_loci)])
    assert len(paths) == 6 and all(os.path.isfile(p) and os.path.isfile(p + ".bai") for p in paths)
    off = np.concatenate([[0], np.cumsum(2 * b.contig_len.astype(np.int64))])
    for l in range(b.n_loci):
        L = int(b.contig_len[l])
        for strand in (0, 1):
            text, refs, recs = realign.read_bam(paths[2 * l + strand])
            assert refs == [("ctg1", L)] and text.startswith("@HD\tVN:1.6\tSO:coordinate")
            mapped = [r for r in recs if not r["flag"] & 4]
            assert [r["pos"] for r in mapped] == sorted(r["pos"] for r in mapped) and all(r["flag"] & 4 for r in recs[len(mapped):])
            want = ro.depth[off[l] + strand * L: off[l] + (strand + 1) * L]
            assert (_depth_from_records(recs, L) == want).all()
            n_reads_l = int(b.locus_read_begin[l + 1] - b.locus_read_begin[l])
            assert len({r["qname"] for r in recs}) == n_reads_l                       # every read has at least one line
            for r in mapped:
                op, ln = r["cigar"] & 15, r["cigar"] >> 4
                qlen_cig = int(ln[(op == 0) | (op == 1) | (op == 4)].sum())
                rd = int(r["qname"][4:])
                full = int(ln[(op == 0) | (op == 1) | (op == 4) | (op == 5)].sum())
                assert full == int(b.read_len[rd])                                     # clips + aligned query = read length
                if r["flag"] & 0x100:
                    assert r["seq"] == "" and (op != 4).all()                          # secondary: SEQ '*', hard clips
                else:
                    assert len(r["seq"]) == qlen_cig
                    assert ((op != 5).all() if not r["flag"] & 0x800 else (op != 4).all())
                assert {"NM", "ms", "AS", "nn", "tp", "cm", "s1", "de"} <= set(r["tags"]) and 0 <= r["mapq"] <= 60
                assert r["tags"]["tp"] == ("S" if r["flag"] & 0x100 else r["tags"]["tp"]) and ("s2" in r["tags"]) == (not r["flag"] & 0x100)
            # SA:Z: every non-secondary line of a read with several of them lists the others (minimap2's abbreviated CIGAR)
            by_name = {}
            for r in mapped:
                if not r["flag"] & 0x100:
                    by_name.setdefault(r["qname"], []).append(r)
            for name, rs_ in by_name.items():
                for r in rs_:
                    assert ("SA" in r["tags"]) == (len(rs_) > 1)
                    if len(rs_) > 1:
                        ents = [e.split(",") for e in r["tags"]["SA"].rstrip(";").split(";")]
                        others = [o for o in rs_ if o is not r]
                        assert sorted((int(e[1]) - 1, e[2], int(e[4]), int(e[5])) for e in ents) == \
                            sorted((o["pos"], "-" if o["flag"] & 0x10 else "+", o["mapq"], o["tags"]["NM"]) for o in others)
                        assert all(e[0] == "ctg1" for e in ents)
                        for e in ents:     # clips + M + I of the abbreviated CIGAR = read length
                            parts = re.findall(r"(\d+)([SMID])", e[3])
                            assert sum(int(n) for n, op in parts if op in "SMI") == int(b.read_len[int(name[4:])])
            # primary SEQ is the read (reverse-complemented for reverse hits)
            prim = next(r for r in mapped if not r["flag"] & 0x900)
            rd = int(prim["qname"][4:])
            codes = b.unpack(int(b.read_off[rd]), int(b.read_len[rd]))
            fw = "".join("ACGTN"[c] for c in codes)
            rc = "".join("TGCAN"[c] for c in codes[::-1])
            assert prim["seq"] == (rc if prim["flag"] & 0x10 else fw)
            # the index serves window queries on the written file
            f = gather.BamFile(paths[2 * l + strand])
            mid = L // 2
            want_names = sorted({r["qname"] for r in mapped if r["pos"] < mid + 50 and r["pos"] + int((r["cigar"] >> 4)[np.isin(r["cigar"] & 15, (0, 2, 3, 7, 8))].sum()) > mid})
            assert sorted(set(f.fetch("ctg1", mid, mid + 50))) == want_names
            f.close()
