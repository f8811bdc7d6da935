#!/usr/bin/env python
"""bench.py — stage-4 allele-frequency throughput (candidate loci/s) on synthetic batches.

  python bench.py --gpus N --steps K --warmup W     this repo's CUDA path, one rank per GPU (torchrun for N > 1)
  python bench.py --impl reference ...              the CPU path timed on the box's host cores: the reference's own
                                                    minimap2 + samtools + Python AF when the binaries exist (PATH or
                                                    baseline/_ref/bin), else the CPU restatement in oracle/

Workload (BASELINE.json, north_star): config 4 `ont_30k_30x` — 30 000 candidate loci at 30x ONT, one job for every N
(strong scaling).  The job is partitioned by locus with the product's own partitioner (stage4.partition_costs, LPT on
read bases + contig length); rank r generates and runs its shard, results are gathered to rank 0 in locus order and
hashed (`outputs_sha1`: the same digest at every N).  No data-path collective: NCCL only carries the barrier, the time
reduction and the 40-byte-per-locus result gather.  TELR_BENCH_CONFIG selects another configuration (profiles/ keeps the
lines of configs 2, 3 and 5).

A "step" is one pass of the whole stage-4 body over the job: sketch -> index/seed/chain -> base-level DP -> depth ->
medians -> AF.  `value` is measured with the batch resident in HBM (telr_af_run_device); `e2e` goes through the
host-buffer C ABI call telr_af_run on the caller's pageable numpy arrays (H2D of the packed batch, overlapped chunk by
chunk with the kernels, and D2H of the results inside the timed region); `e2e_pinned` is the same call on pinned buffers.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = os.environ.get("TELR_BENCH_CONFIG", "ont_30k_30x")     # BASELINE.json configs[3], the north_star's target batch
METRIC = "candidate loci/sec (stage-4 AF)"                         # GCUPS of the base-level DP is reported in `gcups` / `roofline`
DTYPE = "int16x2 (integer DP), fp32 chain penalty, fp64 AF"
OPS_PER_CELL = 30.0        # integer lane-ops per DP cell of the two-piece affine recurrence with traceback
DRAM_BYTES_PER_CELL = float(os.environ.get("TELR_DRAM_B_PER_CELL", "1.47"))   # dram bytes of the DP kernels / DP cells, ncu --set full capture (profiles/)
BUDGET_S = float(os.environ.get("TELR_BENCH_BUDGET_S", "420"))    # the e2e legs shrink their step count to keep the whole run inside this


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arms
def find_real_tools():
    """minimap2 + samtools on PATH or under baseline/_ref/bin (a driver-provided install); None when either is missing."""
    extra = os.path.join(ROOT, "baseline", "_ref", "bin")
    path = os.environ.get("PATH", "") + os.pathsep + extra
    mm, st = shutil.which("minimap2", path=path), shutil.which("samtools", path=path)
    return (mm, st) if mm and st else None


def cpu_sample_loci(cores):
    """>= 8 loci per host thread so that the dynamic OpenMP schedule (or the process pool) is balanced."""
    return int(os.environ.get("TELR_REF_SAMPLE_LOCI", str(8 * max(cores, 1))))


def run_cpu_sample(batch, cores, tools):
    """One pass of the CPU path over `batch`; returns (seconds, dp_cells or None, kind)."""
    if tools:
        from baseline import ref_tools
        t0 = time.time()
        ref_tools.run(batch, tools[0], tools[1], threads=cores)
        return time.time() - t0, None, "reference"
    from tests import orc
    t0 = time.time()
    r = orc.af_run(batch, threads=cores, want_depth=False, want_aln=False)
    return time.time() - t0, int(r.c.dp_cells), "port"


def cpu_baseline(total_loci):
    """Bounded sample of the same workload on the host cores (the oracle is the checker/baseline here, never the product)."""
    from telr_b200 import synth
    cores = os.cpu_count() or 1
    tools = find_real_tools()
    n = min(total_loci, cpu_sample_loci(cores))
    b = synth.generate(WORKLOAD, 0, n)
    dt, cells, kind = run_cpu_sample(b, cores, tools)
    d = {"value": n / dt, "unit": "loci/s", "cores": cores, "kind": kind,
         "sample": f"first {n} loci of {WORKLOAD}, {dt:.1f} s" + ("" if tools else "; CPU restatement of minimap2 2.22 + samtools depth + TELR AF (oracle/), not the real binaries")}
    if cells is not None:
        d["gcups"] = cells / dt / 1e9
    return d


def run_reference(args):
    """CPU arm (rank 0 only): each step is a bounded sample of the workload on all host threads."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from telr_b200 import synth
    cores = os.cpu_count() or 1
    tools = find_real_tools()
    sample = min(synth.CONFIGS[WORKLOAD]["n_loci"], cpu_sample_loci(cores))
    b = synth.generate(WORKLOAD, 0, sample)
    for _ in range(1 if args.warmup > 0 else 0):
        run_cpu_sample(b, cores, tools)
    t0 = time.time()
    cells, kind = 0, "port"
    for _ in range(args.steps):
        _, c, kind = run_cpu_sample(b, cores, tools)
        cells += c or 0
    dt = time.time() - t0
    v = sample * args.steps / dt
    what = ("the reference's own path: minimap2 -a -x <preset> + samtools view/sort/index + samtools depth -aa + TELR AF arithmetic" if tools else
            "CPU restatement of minimap2 2.22 + samtools depth + TELR AF (oracle/), not the real binaries (absent from this image)")
    cb = {"value": v, "unit": "loci/s", "cores": cores, "kind": kind, "sample": f"{sample} loci of {WORKLOAD} per step; {what}"}
    if cells:
        cb["gcups"] = cells / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "loci/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_loci_per_step": sample, "preset": synth.CONFIGS[WORKLOAD]["preset"]},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def host_pack_rate():
    """ASCII -> packed 2-bit + N mask on the host (what precedes telr_af_run in get_af), measured on a 64-Mbase sample."""
    from telr_b200 import lib
    from telr_b200.batch import pack_sequences
    rng = np.random.default_rng(1)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 1 << 20)) for _ in range(8)] * 8
    t0 = time.perf_counter()
    pack_sequences(seqs, lib.lib())
    dt = time.perf_counter() - t0
    return len(seqs) * (1 << 20) / dt


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--loci", type=int, default=int(os.environ.get("TELR_BENCH_LOCI", "0")), help="loci of the whole job (0 = the configuration's full size)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    t_start = time.time()
    import torch
    import torch.distributed as dist
    from telr_b200 import lib, stage4, synth
    from telr_b200.batch import CBatch, CResult

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
    total_loci = args.loci or synth.CONFIGS[WORKLOAD]["n_loci"]

    # ---- partition the job by locus (the product's partitioner) ----
    # rank r first generates the contiguous slice r of the job to learn its loci's costs; the cost vector is all-gathered,
    # every rank runs the same LPT partition, and generates the shard it owns (any subset of loci generates identically)
    t0 = time.time()
    s0, s1 = total_loci * rank // world, total_loci * (rank + 1) // world
    slice_b = synth.generate(WORKLOAD, s0, s1 - s0, total_loci=total_loci)
    cost = stage4.locus_costs(slice_b)
    if world > 1:
        pad = (total_loci + world - 1) // world + 1
        mine = torch.zeros(pad, dtype=torch.int64, device=dev)
        mine[: len(cost)] = torch.from_numpy(cost).to(dev)
        allc = [torch.zeros(pad, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allc, mine)
        cost_all = np.concatenate([allc[r].cpu().numpy()[: total_loci * (r + 1) // world - total_loci * r // world] for r in range(world)])
        shards = stage4.partition_costs(cost_all, world)
        batch = synth.generate(WORKLOAD, loci=shards[rank], total_loci=total_loci)
        del slice_b
    else:
        shards = [list(range(total_loci))]
        batch = slice_b
    my_loci = np.asarray(shards[rank], np.int64)
    gen_s = time.time() - t0

    names = ["seq2", "nmask", "read_off", "read_len", "read_hash", "locus_read_begin", "contig_off", "contig_len", "te_start", "te_end"]

    def to_t(a):
        return torch.from_numpy(a.view(np.int32)) if a.dtype == np.uint32 else torch.from_numpy(a)

    ctx = lib.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    dten = {n: to_t(getattr(batch, n)).to(dev) for n in names}
    cov_d = torch.zeros((max(batch.n_loci, 1), 8), dtype=torch.int32, device=dev)
    af_d = torch.zeros(max(batch.n_loci, 1), dtype=torch.float64, device=dev)
    hd = (batch.preset, batch.flank_len, batch.flank_off, batch.te_len, batch.te_off, batch.n_loci, batch.n_reads, batch.n_bases)
    cb = CBatch(*hd, *[dten[n].data_ptr() for n in names])
    cres = CResult()
    cres.cov2x, cres.af = cov_d.data_ptr(), af_d.data_ptr()
    # result gather (N > 1): every rank contributes a padded [max shard, 10] float64 block; rank 0 scatters the rows into locus order
    max_shard = max(len(s) for s in shards)
    out_all = torch.zeros((total_loci, 10), dtype=torch.float64, device=dev) if rank == 0 else None
    idx_all = [torch.from_numpy(np.asarray(s, np.int64)).to(dev) for s in shards] if rank == 0 else None

    def gather(cov_t, af_t):
        """cov2x/af of this rank's shard -> rank 0, rows in vcf_parsed (locus) order.  40 bytes per locus; no data-path collective."""
        blk = torch.zeros((max_shard, 10), dtype=torch.float64, device=dev)
        n = batch.n_loci
        if n:
            blk[:n, :8] = cov_t[:n].to(torch.float64)
            blk[:n, 8] = torch.nan_to_num(af_t[:n], nan=-1.0)
            blk[:n, 9] = torch.isnan(af_t[:n]).to(torch.float64)
        if world > 1:
            lst = [torch.zeros_like(blk) for _ in range(world)] if rank == 0 else None
            dist.gather(blk, lst, dst=0)
        else:
            lst = [blk]
        if rank == 0:
            for r in range(world):
                out_all[idx_all[r]] = lst[r][: len(shards[r])]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between barriers; device time from CUDA events on the library's stream, max over ranks; also every rank's own time."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        own = 0.0
        for _ in range(steps):
            own += fn()
        torch.cuda.synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) / 1e3
        t = torch.tensor([max(dev_s, 0.0), wall], dtype=torch.float64, device=dev)
        per_rank = [own / steps * 1e3]
        if world > 1:
            allt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allt, torch.tensor([own / steps * 1e3], dtype=torch.float64, device=dev))
            per_rank = [float(x[0]) for x in allt]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1]), per_rank

    stage_ms, cells, counts = {}, [0], {"mz": 0, "blk": 0}

    def step_dev():
        t0 = time.perf_counter()
        if batch.n_loci:
            ctx.run_device(cb, cres)
            for i in range(8):
                stage_ms[i] = stage_ms.get(i, 0.0) + float(cres.ms_stage[i])
            cells[0] += int(cres.dp_cells)
            counts["mz"] += int(cres.n_minimizers); counts["blk"] += int(cres.n_aln_blocks)
        own = time.perf_counter() - t0          # the shard's own pipeline (run_device returns synchronised)
        gather(cov_d, af_d)
        return own

    W = max(args.warmup, 3)
    for _ in range(W):
        step_dev()
    stage_ms.clear(); cells[0] = 0; counts["mz"] = counts["blk"] = 0
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    dev_s, wall_s, per_rank_ms = timed(step_dev, args.steps)
    launches = ctx.launches - l0
    clocks = sampler.finish()
    torch.cuda.synchronize()
    digest = None
    if rank == 0:
        o = out_all.cpu().numpy()
        digest = hashlib.sha1(np.ascontiguousarray(o[:, :8].astype(np.int32)).tobytes() + np.ascontiguousarray(o[:, 8:]).tobytes()).hexdigest()
    del dten
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C ABI call: pageable numpy arrays (what get_af passes), then pinned ----
    hcb = batch.as_c()
    cov_h = np.zeros((max(batch.n_loci, 1), 8), np.int32)
    af_h = np.zeros(max(batch.n_loci, 1), np.float64)
    hres = CResult()
    hres.cov2x, hres.af = cov_h.ctypes.data, af_h.ctypes.data

    def step_host(cbatch=hcb, res=hres, cov=cov_h, af=af_h):
        t0 = time.perf_counter()
        if batch.n_loci:
            rc = lib.lib().telr_af_run(ctx._h, C.byref(cbatch), C.byref(res))
            if rc != 0:
                raise lib.TelrError(rc, "telr_af_run")
        own = time.perf_counter() - t0
        gather(torch.from_numpy(cov).to(dev), torch.from_numpy(af).to(dev))
        return own

    step_ms = dev_s / args.steps
    def steps_for(frac):
        left = BUDGET_S * frac - (time.time() - t_start)
        return int(max(2, min(args.steps, left / max(step_ms, 1e-3) - 1)))
    k_e2e = steps_for(0.8)
    step_host()
    _, e2e_wall, e2e_per_rank = timed(step_host, k_e2e)
    e2e = total_loci * k_e2e / e2e_wall
    # pinned variant
    pinned = {n: to_t(getattr(batch, n)).pin_memory() for n in names}
    pcb = CBatch(*hd, *[pinned[n].data_ptr() for n in names])
    cov_p = torch.zeros((max(batch.n_loci, 1), 8), dtype=torch.int32).pin_memory()
    af_p = torch.zeros(max(batch.n_loci, 1), dtype=torch.float64).pin_memory()
    pres = CResult()
    pres.cov2x, pres.af = cov_p.data_ptr(), af_p.data_ptr()
    k_pin = steps_for(1.0)
    _, pin_wall, _ = timed(lambda: step_host(pcb, pres, cov_p.numpy(), af_p.numpy()), k_pin)
    e2e_pinned = total_loci * k_pin / pin_wall

    value = total_loci * args.steps / dev_s
    hbm_gbs, sm_max_mhz, peak_kind = measured_peaks()
    # roofline of the dominant kernels (base-level DP): GCUPS of this rank's shard against the integer-ALU peak at the observed SM clock
    align_s = stage_ms.get(3, 0.0) / 1e3
    gcups = cells[0] / align_s / 1e9 if align_s > 0 else 0.0
    f_mhz = clocks["sm_mhz"] or sm_max_mhz
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    peak_gcups = nsm * 64 * f_mhz * 1e6 * 2 / OPS_PER_CELL / 1e9      # alu pipe, 16x2 packed ops: 2 cells per lane-op
    seq_lens = np.concatenate([batch.read_len.astype(np.int64), batch.contig_len.astype(np.int64), batch.contig_len.astype(np.int64)])
    sk_bytes = float(((seq_lens + 3) // 4 + (seq_lens + 7) // 8).sum()) * args.steps + 12.0 * counts["mz"]
    dp_bytes = 8.0 * counts["blk"]
    sk_s, de_s = stage_ms.get(0, 0.0) / 1e3, stage_ms.get(6, 0.0) / 1e3
    h2d, d2h = batch.h2d_bytes(), int(batch.n_loci * 40)
    if world > 1:
        tt = torch.tensor([h2d, d2h, cells[0], batch.n_reads, int(batch.read_len.astype(np.int64).sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(tt)
        h2d, d2h, cells_all, reads_all, bases_all = [int(x) for x in tt.tolist()]
    else:
        cells_all, reads_all, bases_all = cells[0], batch.n_reads, int(batch.read_len.astype(np.int64).sum())
    if rank == 0:
        cpu = None
        if world == 1:
            try:
                cpu = cpu_baseline(total_loci)
            except Exception as ex:       # the oracle is a checker, never a dependency of the product path
                cpu = {"error": str(ex)}
        try:
            pack_rate = host_pack_rate()
        except Exception:
            pack_rate = None
        n_chunks_launch = max(launches, 1)
        line = {
            "metric": METRIC, "value": value, "unit": "loci/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": WORKLOAD, "loci": total_loci, "preset": synth.CONFIGS[WORKLOAD]["preset"], "reads": reads_all, "read_bases": bases_all,
                       "partition": "stage4.partition_costs (LPT on read bases + contig length), results gathered to rank 0 in locus order",
                       "shard_loci": [len(s) for s in shards],
                       "l2": "inputs larger than L2 (packed shard %.0f MB on rank 0)" % (batch.h2d_bytes() / 1e6), "sharding": "by locus, no data-path collective"},
            "outputs_sha1": digest, "per_rank_ms": per_rank_ms,
            "gcups": gcups, "gcups_all_ranks": cells_all / max(args.steps, 1) / (dev_s / args.steps) / 1e9, "dp_cells_per_step": cells_all // max(args.steps, 1),
            "stage_ms_per_step": {k: stage_ms.get(i, 0.0) / args.steps for i, k in enumerate(["sketch", "seed_chain", "plan", "align_dp", "", "", "depth_af"]) if k},
            "roofline": {"kernel": "k_al_* (base-level DP of rank 0's shard)", "bound": "int_alu", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS",
                         "frac": gcups / peak_gcups if peak_gcups else None,
                         "traffic": DRAM_BYTES_PER_CELL * cells[0] / max(args.steps, 1),
                         "note": f"peak = {nsm} SM x 64 lane-ops/clk x {f_mhz} MHz x 2 cells/op / {OPS_PER_CELL:.0f} ops/cell; achieved = DP cells / CUDA-event time of the alignment stage; "
                                 f"traffic = DRAM bytes of the DP kernels per step ({DRAM_BYTES_PER_CELL} B/cell from the ncu capture in profiles/ x this run's cells); HBM peak {hbm_gbs} GB/s ({peak_kind}) applies to sketch/depth"},
            "hbm_stages": {"sketch": {"bound": "hbm", "achieved": sk_bytes / sk_s / 1e9 if sk_s > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
                                      "frac": sk_bytes / sk_s / 1e9 / hbm_gbs if sk_s > 0 else None},
                           "depth_af": {"bound": "hbm", "achieved": dp_bytes / de_s / 1e9 if de_s > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
                                        "frac": dp_bytes / de_s / 1e9 / hbm_gbs if de_s > 0 else None}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "loci/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_wall / k_e2e * 1e3, "steps": k_e2e,
                    "host_buffers": "pageable numpy arrays through telr_af_run (the call get_af makes); H2D overlapped chunk-wise with the kernels", "per_rank_ms": e2e_per_rank},
            "e2e_pinned": {"value": e2e_pinned, "unit": "loci/s", "ms_per_step": pin_wall / k_pin * 1e3, "steps": k_pin},
            "host_prep": {"generate_s": gen_s, "pack_gbases_per_s": pack_rate / 1e9 if pack_rate else None,
                          "pack_ms_per_step_est": bases_all / pack_rate * 1e3 if pack_rate else None,
                          "note": "FASTA parse + 2-bit pack happen before telr_af_run and are outside every timed region (SURVEY 8d); rate = telr_pack_seq on one host thread"},
            "gpu_launches": launches, "clocks": clocks, "wall_ms_per_step": wall_s / args.steps * 1e3,
        }
        del n_chunks_launch
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
