#!/usr/bin/env python
"""bench.py -- read-groups/s and BAQ-HMM GCUPS of the marker-mode scoring hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON
line on rank 0.  A "step" is one pass of the hot path over one batch of synthetic read groups.

  value     whole-job read-groups/s with the batch already resident in HBM when the timed
            region starts (sp_upload once, sp_run_resident per step), all ranks summed;
  e2e       the same metric through the reference-facing C ABI with HOST buffers: sp_submit packs
            into pinned memory, copies host->device, runs, copies results device->host (3 batches
            in flight on 3 streams);
  roofline  the dominant kernel k_hmm against the FP64 pipe (measured live with a register-resident
            DFMA kernel; MEASURED_PEAKS.json has no FP64 entry) -- 41 algorithmic flop per band cell
            (SURVEY.md 8(d)); HBM figures are reported beside it to show the kernel is not
            bandwidth-bound;
  cpu_baseline  the CPU oracle (the reference's own marker-path sources + restated probaln_glocal)
            timed on this box's host cores on a bounded sample of the same workload.

Multi-GPU: read groups are independent and are sharded by query-name range across ranks (no
collective on the data path; torch.distributed is only used for the barrier and the max-over-ranks
time).  Scaling is weak: every rank processes its own `--groups` per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# three slots x five streams per context: enough hardware queues that they do not alias (must be in
# the environment before the first CUDA call of the process, which is torch's)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: NCCL's debug output goes to stderr (its version banner is a
# plain printf, see _claim_stdout)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

FLOP_PER_CELL = 41  # 19 forward + 18 backward + 4 MAP


_JSON_FD = None


def _claim_stdout():
    """From here on file descriptor 1 points at stderr, so nothing a library prints (NCCL's version banner,
    torchrun notices) can land on stdout; emit_line() writes the one JSON line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=8192, help="read groups per step and per GPU")
    ap.add_argument("--locus-len", type=int, default=150_000_000, help="haplotype length (config 3: 150 Mb)")
    ap.add_argument("--pool", type=int, default=3, help="distinct pre-generated batches cycled through the steps")
    ap.add_argument("--cpu-sample", type=int, default=0, help="read groups of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preset", default="hifi", choices=["hifi", "ont", "stress"],
                    help="workload family: hifi = BASELINE configs[2] (the bench line); ont / stress = configs[1] / [4], "
                         "for the profiles only")
    return ap.parse_args()


def workload_name(args):
    if args.preset == "ont":
        return (f"ont (BASELINE configs[1] shape): synthetic diploid 2x{args.locus_len / 1e6:g} Mb, simulated ONT reads "
                f"N(30 kb, 8 kb), primary + 1 secondary, --ont preset; step = {args.groups} read groups per GPU")
    if args.preset == "stress":
        return (f"stress (BASELINE configs[4] shape): near-identical repeat copies, up to 8 secondaries per read, "
                f"homopolymer-rich, --hifi preset; step = {args.groups} read groups per GPU")
    return (f"chr-hifi-30x (BASELINE configs[2]): synthetic diploid 2x{args.locus_len / 1e6:g} Mb, simulated HiFi "
            f"reads N(15 kb, 2 kb), primary + 1 secondary, --hifi preset; step = {args.groups} read groups per GPU")


def make_synth(args):
    from tools.synth.pysynth import Synth, default_cfg
    return Synth(default_cfg(args.preset, locus_len=args.locus_len, seed=20240603))


def params_preset(args):
    return "ont" if args.preset == "ont" else "hifi"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def count(self, t0, t1=None):
        return sum(1 for t, _ in self.lines if t >= t0 and (t1 is None or t <= t1))

    def stop(self, t0=None, t1=None):
        """Summary over the samples that arrived in [t0, t1] (the timed regions; default: all)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for t, ln in self.lines:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(synth, batches, preset, threads, keep=None):
    """Time the CPU oracle over `batches` (list of FlatBatch) with a thread pool over read groups
    (the reference's own parallelism: task-parallel over read groups, tpool.c).  Returns
    (groups, cells, seconds, kind).  keep: a dict that receives the oracle's scores and selected
    indices in group order (used to check the GPU results of the same groups, never timed work)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pyoracle
    kinds = pyoracle.available_kinds()
    if not kinds:
        try:
            pyoracle.build()
        except Exception:
            pass
        kinds = pyoracle.available_kinds()
    if not kinds:
        raise RuntimeError("no CPU reference library: oracle/_ref was not built here and /root/reference is not mounted")
    kind = kinds[0]
    ref = pyoracle.make_refseq(synth.names, [synth.contig_ptr(i) for i in range(synth.n_contigs)], synth.lens)
    params = pyoracle.preset_params(preset)
    chunks = []
    for b in batches:
        per = max(1, (b.n_groups + threads * 4 - 1) // (threads * 4))
        for g0 in range(0, b.n_groups, per):
            chunks.append(b.group_slice(g0, min(b.n_groups, g0 + per)))
    pyoracle.run(chunks[0].group_slice(0, 1), params, ref, kind=kind)  # load + warm

    def work(c):
        r = pyoracle.run(c, params, ref, kind=kind, seed=None)
        h = r["hmm"]
        return int(h[:, 4].astype(np.int64).sum() + (h[:, 5].astype(np.int64) << 31).sum()), r["scores"], r["groups"][:, 0]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, chunks))
    dt = time.perf_counter() - t0
    cells = sum(p[0] for p in parts)
    if keep is not None:
        keep["scores"] = np.concatenate([p[1] for p in parts])
        keep["best"] = np.concatenate([p[2] for p in parts])
    return sum(b.n_groups for b in batches), cells, dt, kind


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    synth = make_synth(args)
    # bounded sample per step so that the whole run stays within a few minutes
    per_step = args.cpu_sample or min(args.groups, 128 * threads)
    n_steps = args.warmup + args.steps
    batches = [synth.generate(i * per_step, per_step) for i in range(min(n_steps, 3))]
    times, groups, cells = [], 0, 0
    kind = "port"
    for i in range(n_steps):
        try:
            g, c, dt, kind = cpu_reference_run(synth, [batches[i % len(batches)]], params_preset(args), threads)
        except RuntimeError as e:
            emit_line({"impl": "reference", "unavailable": str(e)})
            return 0
        if i >= args.warmup:
            times.append(dt)
            groups += g
            cells += c
    total = sum(times)
    val = groups / total
    line = {
        "impl": "reference", "metric": "read_groups_per_sec", "value": val, "unit": "read-groups/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": f"{per_step} read groups per step on the host CPU"},
        "gcups": cells / total / 1e9,
        "cpu_baseline": {"value": val, "unit": "read-groups/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} read groups x {args.steps} steps, thread pool over read groups",
                         "gcups": cells / total / 1e9},
        "e2e": {"value": val, "unit": "read-groups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def main():
    args = parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device; this framework has no CPU path\n")
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import secphase_b200
    from tools.parity import encode_reference

    t_setup = time.perf_counter()
    synth = make_synth(args)
    codes, off = encode_reference(synth)
    eng = secphase_b200.Secphase(params_preset(args), device=local)
    eng.set_reference_codes(codes, off)
    # shard by query-name range: rank r owns groups [r*span, (r+1)*span)
    span = args.groups * args.pool
    # the pools of every batch live in page-locked host memory, as a reader thread decoding straight
    # into sp_host_alloc'ed buffers would leave them (include/secphase_b200.h)
    batches = [secphase_b200.pin_batch(synth.generate(rank * span + i * args.groups, args.groups))
               for i in range(args.pool)]
    setup_s = time.perf_counter() - t_setup

    n_slots = min(3, args.pool)

    # ---- device-resident arm (value) -------------------------------------------------------
    for s in range(n_slots):
        eng.upload(batches[s], slot=s)

    def resident_steps(n):
        stats = []
        inflight = []
        for i in range(n):
            s = i % n_slots
            if len(inflight) == n_slots:
                stats.append(eng.wait(inflight.pop(0), copy=False))
            eng.run_resident(s)
            inflight.append(s)
        while inflight:
            stats.append(eng.wait(inflight.pop(0), copy=False))
        return stats

    def device_span_ms(n):
        """CUDA-event time from the mark to the end of the last batch on any slot used by n steps."""
        return max(eng.elapsed_since_mark(s) for s in range(min(n, n_slots)))

    # nvidia-smi needs a second or two before its first sample: start it ahead of the warm-up and keep only
    # the samples that arrive inside the measured phases
    sampler = ClockSampler(local)
    sampler.start()
    resident_steps(args.warmup)
    barrier()
    t_meas0 = time.perf_counter()
    eng.mark()
    t0 = time.perf_counter()
    stats = resident_steps(args.steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev_s = device_span_ms(args.steps) * 1e-3
    dev_wall_s = t1 - t0
    barrier()

    # ---- end-to-end arm (host buffers, H2D + D2H inside the timed region) ------------------
    def e2e_steps(n):
        out = []
        inflight = []
        for i in range(n):
            s = i % n_slots
            if len(inflight) == n_slots:
                out.append(eng.wait(inflight.pop(0), copy=True))
            eng.submit(batches[i % len(batches)], slot=s)
            inflight.append(s)
        while inflight:
            out.append(eng.wait(inflight.pop(0), copy=True))
        return out

    e2e_steps(max(args.warmup, 3))
    barrier()
    # the end-to-end arm is a host-visible quantity (pack + H2D + kernels + D2H + result tables in
    # host memory), so it is timed on the host clock around the synchronised region; the device-side
    # span of the same steps is reported next to it
    eng.mark()
    t2 = time.perf_counter()
    estats = e2e_steps(args.steps)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    e2e_dev_s = device_span_ms(args.steps) * 1e-3
    barrier()
    e2e_s = t3 - t2

    # ---- the dominant kernel alone: one batch in flight, nothing overlapped, so that the CUDA
    # events around the HMM launches (fork -> class kernels on aux streams -> join, recorded on the
    # slot's stream) time just those kernels
    iso = []
    for i in range(2 + max(3, min(args.steps, 8))):
        eng.run_resident(0)
        r = eng.wait(0, copy=False)
        if i >= 2:
            iso.append(r)
    # the measured phases are short (~1.5 s); if the sampler caught fewer than three readings under load,
    # keep the same resident workload running (untimed) until it has
    t_load1 = time.perf_counter()
    while sampler.proc and sampler.count(t_meas0) < 3 and time.perf_counter() - t_load1 < 6.0:
        resident_steps(n_slots)
    clocks = sampler.stop(t_meas0, time.perf_counter())
    # ---- roofline denominators, measured live ------------------------------------------------
    dfma_ops, _ = eng.fp64_peak(0)
    dadd_ops, _ = eng.fp64_peak(1)

    if world > 1:
        t = torch.tensor([dev_s, e2e_s, dev_wall_s, e2e_dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, dev_wall_s, e2e_dev_s = (float(x) for x in t)

    groups_total = args.groups * args.steps * world
    cells_step = float(np.mean([s["hmm_cells"] for s in stats]))
    hmm_ms = float(np.mean([s["ms_hmm"] for s in iso]))
    iso_cells = float(np.mean([s["hmm_cells"] for s in iso]))
    stage_ms = np.mean([s["ms_stage"] for s in iso], axis=0).tolist()
    launches = int(sum(s["gpu_launches"] for s in stats))
    value = groups_total / dev_s
    gcups_kernel = iso_cells / (hmm_ms * 1e-3) / 1e9
    gcups_job = cells_step * args.steps * world / dev_s / 1e9
    achieved_tflops = iso_cells * FLOP_PER_CELL / (hmm_ms * 1e-3) / 1e12
    peak_tflops = 2.0 * dfma_ops / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # Algorithmic HBM bytes of one launch set: reference window + packed query in
    # (~0.05 B/cell) plus the per-row scaling factors, which must survive from the
    # forward to the backward sweep: 8 B out + 8 B back in per query row (~0.35 B/cell at 41 cells/row)
    alg_bytes = 0.4 * iso_cells
    # measured DRAM traffic of the same launch set from the committed ncu --set full capture,
    # scaled by band cells to this run's batch
    traffic = None
    traffic_src = None
    try:
        if args.preset != "hifi":
            raise KeyError("the capture is of the HiFi workload")
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_k_hmm2.json")))
        traffic = tr["dram_bytes_total"] * iso_cells / tr["band_cells"]
        traffic_src = tr["source"]
    except Exception:
        traffic_src = None
    line = {
        "metric": "read_groups_per_sec", "value": value, "unit": "read-groups/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "groups_per_step_per_gpu": args.groups, "slots_in_flight": n_slots,
                   "l2": "per-step input (~%.0f MB) exceeds the 126 MB L2; no explicit flush" % (stats[0]["h2d_bytes"] / 1e6 if stats[0]["h2d_bytes"] else estats[0]["h2d_bytes"] / 1e6),
                   "parallelism": f"read groups sharded by qname range over {world} GPU(s), no collective",
                   "sm_partition": dict(zip(("on", "int_sms", "hmm_sms"), eng.sm_partition()))},
        "gcups": gcups_job, "gcups_kernel": gcups_kernel,
        "hmm_instances_per_step": float(np.mean([s["hmm_instances"] for s in stats])),
        "band_cells_per_step": cells_step,
        "e2e": {"value": groups_total / e2e_s, "unit": "read-groups/s",
                "h2d_bytes_per_step": int(np.mean([s["h2d_bytes"] for s in estats])),
                "d2h_bytes_per_step": int(np.mean([s["d2h_bytes"] for s in estats])),
                "ms_per_step": 1e3 * e2e_s / args.steps, "device_ms_per_step": 1e3 * e2e_dev_s / args.steps,
                "timing": "host clock around the synchronised region (host packing and result tables are part of "
                          "the call); device_ms_per_step = CUDA-event span of the same steps"},
        "timing": "CUDA events: mark on an idle device -> end of the last batch on each slot stream, max over slots "
                  "and ranks (host wall clock of the same region: %.3f ms/step)" % (1e3 * dev_wall_s / args.steps),
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "k_hmm2 (all band-class launches of one step, forked on aux streams)", "bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops if peak_tflops else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch set (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                     "peak_source": "measured live: register-resident DFMA kernel, 2 flop/instr (MEASURED_PEAKS.json has no FP64 entry)",
                     "issue_slot_frac": iso_cells * FLOP_PER_CELL / (hmm_ms * 1e-3) / dadd_ops if dadd_ops else None,
                     "dadd_dmul_ops_per_s": dadd_ops, "kernel_ms_per_step": hmm_ms,
                     "kernel_share_of_step": hmm_ms / float(np.mean([s["ms_total"] for s in iso])),
                     "timing": "CUDA events on the slot stream around the HMM launches, one batch in flight",
                     "hbm": {"algorithmic_gbs": alg_bytes / (hmm_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "stage_ms_isolated": dict(zip(["h2d", "walk", "group", "emit_sort", "hmm", "score", "d2h"], stage_ms[:7])),
        "setup_s": setup_s,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:  # the CPU baseline is an N=1 leg only
        threads = os.cpu_count() or 1
        # bounded sample: whole batches of the same workload until ~12 s of CPU work are done
        n_cpu = args.cpu_sample or min(args.groups, 1024 * threads)
        g = c = 0
        dt = 0.0
        kind = "port"
        i = 0
        parity = None
        try:
            cpu_reference_run(synth, [batches[0].group_slice(0, 1)], params_preset(args), 1)
            have_cpu = True
        except RuntimeError as e:
            have_cpu = False
            line["cpu_baseline"] = {"unavailable": str(e)}
        while have_cpu and dt < 12.0 and i < 8:
            sample = batches[i % len(batches)].group_slice(0, n_cpu)
            keep = {} if i == 0 else None
            g1, c1, dt1, kind = cpu_reference_run(synth, [sample], params_preset(args), threads, keep=keep)
            g, c, dt, i = g + g1, c + c1, dt + dt1, i + 1
            if keep is not None:
                # the same read groups through the GPU path: alignment scores (bit patterns) and the selected
                # alignment must equal the reference's -- at the benchmark's own batch size
                gpu = eng.run(sample, slot=0)
                same_s = bool(np.array_equal(gpu["scores"].view(np.int64), keep["scores"].view(np.int64)))
                # (stress: several secondaries can tie, and the reference then draws from an unseeded rand())
                same_b = bool(np.array_equal(gpu["groups"][:, 0], keep["best"])) if args.preset != "stress" else None
                parity = {"read_groups_checked": int(sample.n_groups), "alignments_checked": int(sample.n_alns),
                          "scores_bit_exact": same_s, "selected_alignment_equal": same_b,
                          "selected_secondaries": int((gpu["groups"][:, 0] != gpu["groups"][:, 1]).sum())}
                if not same_s or same_b is False:
                    sys.stderr.write("bench.py: GPU results differ from the CPU reference on the checked batch\n")
        if have_cpu:
            line["cpu_baseline"] = {"value": g / dt, "unit": "read-groups/s", "cores": threads, "kind": kind,
                                    "sample": f"{i} x the first {n_cpu} read groups of a step ({g} groups, {dt:.1f} s), "
                                              "thread pool over read groups",
                                    "gcups": c / dt / 1e9, "seconds": dt}
        line["parity_vs_cpu_reference"] = parity
    if rank == 0:
        emit_line(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
