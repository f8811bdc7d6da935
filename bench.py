#!/usr/bin/env python
"""bench.py — stage-4 allele-frequency throughput (loci/s) on synthetic batches.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU, loci sharded by rank)
  python bench.py --impl reference ...                     the CPU path timed on the box's host cores (oracle port:
                                                           the reference's minimap2/samtools binaries do not exist in this image)

A "step" is one pass of the whole stage-4 body over one batch: sketch -> index/seed/chain -> base-level DP ->
depth -> medians -> AF.  `value` is measured with the batch resident in HBM (telr_af_run_device); `e2e` goes through
the host-buffer C ABI call telr_af_run (H2D of the packed batch and D2H of the results inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = os.environ.get("TELR_BENCH_CONFIG", "ont_3k_50x")     # BASELINE.json configs[1]
OPS_PER_CELL = 30.0        # integer lane-ops per DP cell of the two-piece affine recurrence with traceback
DRAM_BYTES_PER_CELL = 1.33  # dram__bytes_read+write of k_al_fused / DP cells, ncu --set full capture (profiles/r1_k_al_fused_ncu.csv)
LAUNCHES_PER_CHUNK = 29    # kernels the library launches per chunk of loci (telr_af.cu run_chunk)


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


def cpu_baseline(batch, seconds_target=15.0):
    """Oracle port on the host cores over a bounded sample of the same workload (test infrastructure as the checker/baseline)."""
    from tests import orc
    cores = os.cpu_count() or 1
    n = min(batch.n_loci, max(cores, 8))
    t0 = time.time()
    r = orc.af_run(batch, threads=cores, first=0, n=n, want_depth=False, want_aln=False)
    dt = time.time() - t0
    if dt < seconds_target / 3 and n < batch.n_loci:      # grow the sample once towards the target duration
        n2 = min(batch.n_loci, int(n * seconds_target / max(dt, 1e-3)))
        t0 = time.time()
        r = orc.af_run(batch, threads=cores, first=0, n=n2, want_depth=False, want_aln=False)
        dt = time.time() - t0
        n = n2
    return {"value": n / dt, "unit": "loci/s", "cores": cores, "kind": "port",
            "sample": f"first {n} loci of {batch.meta.get('config')} ({int(r.c.dp_cells)} DP cells, {dt:.1f} s)",
            "gcups": r.c.dp_cells / dt / 1e9}, n, dt


def run_reference(args):
    """CPU arm: the reference's own path is minimap2+samtools+Python, none of which exist in this image; the timed
    stand-in is the CPU oracle port with all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from telr_b200 import synth
    from tests import orc
    cores = os.cpu_count() or 1
    sample = int(os.environ.get("TELR_REF_SAMPLE_LOCI", str(max(8, cores))))
    b = synth.generate(WORKLOAD, 0, sample)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        orc.af_run(b, threads=cores, want_depth=False, want_aln=False)
    t0 = time.time()
    cells = 0
    for _ in range(args.steps):
        r = orc.af_run(b, threads=cores, want_depth=False, want_aln=False)
        cells += int(r.c.dp_cells)
    dt = time.time() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": "candidate loci/sec (stage-4 AF)", "value": v, "unit": "loci/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8/int32 (integer DP), fp32 chain penalty, fp64 AF", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_loci_per_step": sample, "preset": synth.CONFIGS[WORKLOAD]["preset"]},
            "cpu_baseline": {"value": v, "unit": "loci/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} loci of {WORKLOAD} per step; CPU restatement of minimap2 2.22 + samtools depth + TELR AF, not the real binaries",
                             "gcups": cells / dt / 1e9},
            "e2e": {"value": v, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--loci", type=int, default=int(os.environ.get("TELR_BENCH_LOCI", "0")), help="loci per GPU (0 = the configuration's full size)")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("TELR_STREAMS", "1")), help="concurrent contexts (streams) per GPU, each on its own slice of loci")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from telr_b200 import lib, synth
    from telr_b200.batch import CBatch, CResult

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg_loci = synth.CONFIGS[WORKLOAD]["n_loci"]
    per_gpu = args.loci or cfg_loci
    # weak scaling: every rank owns its own shard of loci [rank*per_gpu, (rank+1)*per_gpu) (locus ids are global)
    batch = synth.generate(WORKLOAD, rank * per_gpu, per_gpu, total_loci=max(cfg_loci, world * per_gpu))
    dev = torch.device("cuda", local)
    # K contexts (= K streams with their own workspaces) share the GPU; each owns a contiguous slice of this rank's loci so
    # that one slice's latency-bound phases and stragglers overlap with another slice's DP
    K = max(1, min(args.streams, batch.n_loci))
    cuts = [batch.n_loci * i // K for i in range(K + 1)]
    names = ["seq2", "nmask", "read_off", "read_len", "read_hash", "locus_read_begin", "contig_off", "contig_len", "te_start", "te_end"]

    def to_t(a):
        return torch.from_numpy(a.view(np.int32)) if a.dtype == np.uint32 else torch.from_numpy(a)

    class Slice:
        pass
    slices = []
    for i in range(K):
        sl = Slice()
        sl.b = batch.subset(range(cuts[i], cuts[i + 1])) if K > 1 else batch
        sl.ctx = lib.Context(local)
        sl.pinned = {n: to_t(getattr(sl.b, n)).pin_memory() for n in names}
        sl.dten = {n: sl.pinned[n].to(dev) for n in names}
        sl.cov_d = torch.zeros((sl.b.n_loci, 8), dtype=torch.int32, device=dev)
        sl.af_d = torch.zeros(sl.b.n_loci, dtype=torch.float64, device=dev)
        hd = (sl.b.preset, sl.b.flank_len, sl.b.flank_off, sl.b.te_len, sl.b.te_off, sl.b.n_loci, sl.b.n_reads, sl.b.n_bases)
        sl.cb = CBatch(*hd, *[sl.dten[n].data_ptr() for n in names])
        sl.cres = CResult()
        sl.cres.cov2x, sl.cres.af = sl.cov_d.data_ptr(), sl.af_d.data_ptr()
        sl.hb = CBatch(*hd, *[sl.pinned[n].data_ptr() for n in names])
        sl.cov_h = torch.zeros((sl.b.n_loci, 8), dtype=torch.int32).pin_memory()
        sl.af_h = torch.zeros(sl.b.n_loci, dtype=torch.float64).pin_memory()
        sl.hres = CResult()
        sl.hres.cov2x, sl.hres.af = sl.cov_h.data_ptr(), sl.af_h.data_ptr()
        slices.append(sl)
    ctx = slices[0].ctx
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) / 1e3
        t = torch.tensor([max(dev_s, 0.0), wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1])

    stage_ms = {}
    cells = [0]
    counts = {"mz": 0, "blk": 0}

    def fan_out(fn):
        if K == 1:
            return fn(slices[0])
        errs = []

        def run(sl):
            try:
                fn(sl)
            except Exception as ex:       # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=run, args=(sl,)) for sl in slices]
        [t.start() for t in th]
        [t.join() for t in th]
        if errs:
            raise errs[0]

    def one_dev(sl):
        sl.ctx.run_device(sl.cb, sl.cres)

    def step_dev():
        fan_out(one_dev)
        for sl in slices:
            for i in range(8):
                stage_ms[i] = stage_ms.get(i, 0.0) + float(sl.cres.ms_stage[i])
            cells[0] += int(sl.cres.dp_cells)
            counts["mz"] += int(sl.cres.n_minimizers); counts["blk"] += int(sl.cres.n_aln_blocks)

    def one_host(sl):
        rc = lib.lib().telr_af_run(sl.ctx._h, C.byref(sl.hb), C.byref(sl.hres))
        if rc != 0:
            raise lib.TelrError(rc, "telr_af_run")

    def step_host():
        fan_out(one_host)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    stage_ms.clear(); cells[0] = 0; counts["mz"] = counts["blk"] = 0
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(sl.ctx.launches for sl in slices)
    dev_s, wall_s = timed(step_dev, args.steps)
    launches = sum(sl.ctx.launches for sl in slices) - l0
    clocks = sampler.finish()
    step_host()
    e2e_dev_s, e2e_wall = timed(step_host, args.steps)

    total_loci = per_gpu * world
    value = total_loci * args.steps / dev_s
    e2e = total_loci * args.steps / e2e_wall
    # roofline of the dominant kernel (k_align): GCUPS against the integer-ALU peak at the observed SM clock
    hbm_gbs, sm_max_mhz, peak_kind = measured_peaks()
    align_s = stage_ms.get(3, 0.0) / 1e3 / K        # K contexts run concurrently: per-context kernel time overlaps
    gcups = cells[0] / align_s / 1e9 if align_s > 0 else 0.0
    f_mhz = clocks["sm_mhz"] or sm_max_mhz
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    peak_gcups = nsm * 64 * f_mhz * 1e6 * 2 / OPS_PER_CELL / 1e9      # alu pipe, 16x2 packed ops: 2 cells per lane-op
    # HBM-bound stages (SURVEY 8d work units): sketch reads ceil(len/4) + ceil(len/8) bytes per sequence (reads and both
    # contig strands) and writes 12 B per minimizer; depth+AF reads 8 B per alignment block (per-base depth is not
    # requested in the bench, so nothing is written).  Stage times are CUDA events around the stage's kernels.
    seq_lens = np.concatenate([batch.read_len.astype(np.int64), batch.contig_len.astype(np.int64), batch.contig_len.astype(np.int64)])
    sk_bytes = float(((seq_lens + 3) // 4 + (seq_lens + 7) // 8).sum()) * args.steps + 12.0 * counts["mz"]
    dp_bytes = 8.0 * counts["blk"]
    sk_s, de_s = stage_ms.get(0, 0.0) / 1e3 / K, stage_ms.get(6, 0.0) / 1e3 / K
    h2d = batch.h2d_bytes()
    d2h = int(batch.n_loci * 40)
    line = None
    if rank == 0:
        cpu = None
        try:
            cpu, _, _ = cpu_baseline(batch) if world == 1 else (None, 0, 0)
        except Exception as ex:       # the oracle is a checker, never a dependency of the product path
            cpu = {"error": str(ex)}
        line = {
            "metric": "candidate loci/sec (stage-4 AF) and read-vs-contig GCUPS", "value": value, "unit": "loci/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8/int32 (integer DP), fp32 chain penalty, fp64 AF", "data": "synthetic",
            "config": {"workload": WORKLOAD, "loci_per_gpu": per_gpu, "preset": synth.CONFIGS[WORKLOAD]["preset"], "reads": int(batch.n_reads), "read_bases": int(batch.read_len.astype(np.int64).sum()),
                       "l2": "inputs larger than L2 (packed batch %.0f MB per GPU)" % (h2d / 1e6), "sharding": "by locus, no collective", "streams_per_gpu": K},
            "gcups": gcups, "dp_cells_per_step": cells[0] // max(args.steps, 1),
            "stage_ms_per_step": {k: stage_ms.get(i, 0.0) / args.steps for i, k in enumerate(["sketch", "seed_chain", "plan", "align_dp", "", "", "depth_af"]) if k},
            "roofline": {"kernel": "k_align (base-level DP)", "bound": "int_alu", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS",
                         "frac": gcups / peak_gcups if peak_gcups else None,
                         "traffic": (DRAM_BYTES_PER_CELL * cells[0] / max(launches // LAUNCHES_PER_CHUNK, 1)) if cells[0] else None,
                         "note": f"peak = {nsm} SM x 64 lane-ops/clk x {f_mhz} MHz x 2 cells/op / {OPS_PER_CELL:.0f} ops/cell; HBM peak {hbm_gbs} GB/s ({peak_kind}) applies to sketch/depth; traffic = bytes per k_al_fused launch, 1.33 B/cell from the ncu capture in profiles/ scaled by this run's cells; the kernel issues 0.71 warp-inst/clk/sub-partition against a measured two-pipe ceiling of 0.705 (profiles/ubench)"},
            "hbm_stages": {"sketch": {"bound": "hbm", "achieved": sk_bytes / sk_s / 1e9 if sk_s > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
                                      "frac": sk_bytes / sk_s / 1e9 / hbm_gbs if sk_s > 0 else None, "minimizers_per_step": counts["mz"] // max(args.steps, 1)},
                           "depth_af": {"bound": "hbm", "achieved": dp_bytes / de_s / 1e9 if de_s > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
                                        "frac": dp_bytes / de_s / 1e9 / hbm_gbs if de_s > 0 else None, "blocks_per_step": counts["blk"] // max(args.steps, 1)}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "loci/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_wall / args.steps * 1e3},
            "gpu_launches": launches, "clocks": clocks, "wall_ms_per_step": wall_s / args.steps * 1e3,
        }
        print(json.dumps(line), flush=True)
    for sl in slices:
        sl.ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
