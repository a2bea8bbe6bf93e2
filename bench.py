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

  per_config  (N=1 only) the other BASELINE configs, each at its own size, measured the same way
            (value / e2e / roofline fraction / stage times) and parity-checked AT SIZE against the
            reference's own code on the host: score bit patterns, selected alignment (the reference's
            get_best_record_index replayed in group order on one rand() stream), the BAQ at every
            marker, final markers, consensus blocks and alignment extents of >= 4096 read groups.

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
    ap.add_argument("--groups", type=int, default=0, help="read groups per step and per GPU (0 = the config's own)")
    ap.add_argument("--locus-len", type=int, default=0, help="haplotype length (0 = the config's own; configs[2]: 150 Mb)")
    ap.add_argument("--pool", type=int, default=0, help="distinct pre-generated batches cycled through the steps (0 = the config's own)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="read groups of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the other BASELINE configs (N=1 default run)")
    ap.add_argument("--parity-groups", type=int, default=4096, help="read groups of the at-size parity check per config")
    ap.add_argument("--hmm", default=None, choices=["fast", "strict"], help="HMM arithmetic (default: the library's)")
    ap.add_argument("--preset", default="hifi", choices=sorted(SPECS),
                    help="workload: hifi = BASELINE configs[2] (the bench line); hifi_small / ont / wg_offsets / stress = "
                         "configs[0] / [1] / [3] / [4] (reported under per_config by the default run)")
    return ap.parse_args()


# The BASELINE.json configs as workloads.  `hifi` (configs[2]) is the bench line; the others are measured the
# same way in `per_config` at N=1.  filler_bp: 'N' contigs placed BEFORE the real ones so that every window
# the kernels fetch lies beyond 2^32 in the device replica (configs[3]: a 2 x 3.1 Gb assembly is 6.2 GB of codes).
SPECS = {
    "hifi": dict(config="configs[2]", title="chr-hifi-30x", synth="hifi", params="hifi", locus_len=150_000_000,
                 groups=8192, pool=3, reads="simulated HiFi reads N(15 kb, 2 kb), primary + 1 secondary, --hifi preset"),
    "hifi_small": dict(config="configs[0]", title="hifi-10k", synth="hifi", params="hifi", locus_len=5_000_000,
                       groups=2048, pool=5, reads="10 240 simulated HiFi reads (~15 kb), primary + 1 secondary, --hifi preset"),
    # large batches: the integer stages are thread-per-alignment / per-group latency chains whose time barely grows with the
    # batch (profiles/r02_stage_ont_batch_size_v13.json: 4096 groups 38 ms, 20 480 groups 68 ms), so ONT wants many
    # alignments in flight
    "ont": dict(config="configs[1]", title="ont-30k", synth="ont", params="ont", locus_len=5_000_000,
                groups=10240, pool=3, reads="30 720 simulated ONT reads N(30 kb, 8 kb) (the config asks for 20k), higher indel rate, primary + 1 secondary, --ont preset"),
    "wg_offsets": dict(config="configs[3]", title="wg-hifi-30x addressing", synth="hifi", params="hifi", locus_len=20_000_000,
                       groups=8192, pool=3, filler_bp=6_160_000_000,
                       reads="HiFi reads on the last two contigs of a 6.2 Gb replica (25 filler contigs of N first: every "
                             "reference window lies beyond 2^32), --hifi preset"),
    "stress": dict(config="configs[4]", title="stress", synth="stress", params="hifi", locus_len=2_000_000,
                   groups=8192, pool=3,
                   # a second parity sample from even more similar copies: there several secondaries share the top
                   # score, so the reference's rand() tie-break (ptAlignment.c:156-170) is compared at size too
                   tie_parity=dict(groups=1024, synth_over=dict(snv_rate=2e-4, indel_rate=4e-5, long_indel_rate=4e-6)),
                   reads="9 near-identical repeat copies, up to 8 secondaries per read, homopolymer-rich (30 %), --hifi preset"),
}


def spec_for(args):
    sp = dict(SPECS[args.preset])
    if args.locus_len:
        sp["locus_len"] = args.locus_len
    if args.groups:
        sp["groups"] = args.groups
    if args.pool:
        sp["pool"] = args.pool
    return sp


def workload_name(sp, steps=None):
    tile = ""
    if steps is not None:
        tile = (f"; {sp['pool']} distinct pre-generated batches cycled over the steps "
                f"(tiling factor {steps / sp['pool']:.1f}x)")
    n_ctg = "2x" if sp["synth"] != "stress" else "9x"
    return (f"{sp['title']} (BASELINE {sp['config']}): synthetic assembly {n_ctg}{sp['locus_len'] / 1e6:g} Mb, {sp['reads']}; "
            f"step = {sp['groups']} read groups per GPU{tile}")


class Workload:
    """Synthetic assembly + read groups of one spec, the encoded replica for the device, and the matching
    view for the CPU oracle (same contig table, filler contigs included)."""

    def __init__(self, sp):
        from tools.parity import encode_reference
        from tools.synth.pysynth import Synth, default_cfg
        self.sp = sp
        self.synth = Synth(default_cfg(sp["synth"], locus_len=sp["locus_len"], seed=20240603, **sp.get("synth_over", {})))
        codes, off = encode_reference(self.synth)
        self.tid_shift = 0
        self.names = list(self.synth.names)
        self.lens = list(self.synth.lens)
        filler = int(sp.get("filler_bp", 0))
        if filler:
            per = 246_400_000  # chromosome-1-sized filler contigs, each well inside BAM's 2^31 position range
            n_fill = max(1, filler // per)
            fl = [filler // n_fill] * n_fill
            self.tid_shift = n_fill
            big = np.empty(sum(fl) + len(codes), np.uint8)
            big[:sum(fl)] = 4
            big[sum(fl):] = codes
            codes = big
            off = np.concatenate([np.cumsum([0] + fl)[:-1], off + sum(fl)]).astype(np.int64)
            self.names = [f"fill{i}" for i in range(n_fill)] + self.names
            self.lens = fl + self.lens
        self.codes, self.off = codes, off
        self._dummy = np.full(64, ord("N"), np.uint8)

    def generate(self, first, n):
        b = self.synth.generate(first, n)
        if self.tid_shift:
            b.tid = (b.tid + self.tid_shift).astype(np.int32)
        return b

    def oracle_refseq(self, pyoracle):
        s = self.synth
        ptrs = [self._dummy.ctypes.data] * self.tid_shift + [s.contig_ptr(i) for i in range(s.n_contigs)]
        return pyoracle.make_refseq(self.names, ptrs, self.lens)


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


def _oracle():
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
    return pyoracle, kinds[0]


_OFF_TABLES = ("markers_pre", "markers_baq", "markers_final")


def _concat_oracle(parts):
    """Results of consecutive group ranges -> one result (offset tables re-based)."""
    out = {}
    for k in ("groups", "scores", "extents", "blocks", "hmm") + _OFF_TABLES:
        out[k] = np.concatenate([p[k] for p in parts])
    for k in ("block_off",) + tuple(t + "_off" for t in _OFF_TABLES):
        base, pieces = 0, [np.zeros(1, np.int64)]
        for p in parts:
            pieces.append(p[k][1:] + base)
            base += int(p[k][-1])
        out[k] = np.concatenate(pieces)
    return out


def cpu_reference_run(wl, batches, preset, threads, keep=None):
    """Time the CPU oracle over `batches` (list of FlatBatch) with a thread pool over read groups
    (the reference's own parallelism: task-parallel over read groups, tpool.c).  Returns
    (groups, cells, seconds, kind).  keep: a dict that receives every table the oracle records, in
    group order (used to check the GPU results of the same groups, never timed work); the selected
    alignment is then the reference's own get_best_record_index replayed over the groups in order on
    one rand() stream seeded like a fresh process -- what a single-worker run of the reference selects
    (the pool's workers share one rand() state, so their own draws interleave arbitrarily)."""
    from concurrent.futures import ThreadPoolExecutor
    pyoracle, kind = _oracle()
    ref = wl.oracle_refseq(pyoracle)
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
        cells = int(h[:, 4].astype(np.int64).sum() + (h[:, 5].astype(np.int64) << 31).sum())
        if keep is None:
            return cells, None
        return cells, r

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, chunks))
    dt = time.perf_counter() - t0
    cells = sum(p[0] for p in parts)
    if keep is not None:
        res = _concat_oracle([p[1] for p in parts])
        flags = np.concatenate([b.flag for b in batches])
        gao = np.concatenate([np.zeros(1, np.int64)] + [
            b.grp_aln_off[1:].astype(np.int64) + sum(x.n_alns for x in batches[:i]) for i, b in enumerate(batches)])
        if kind == "reference":
            res["groups"] = res["groups"].copy()
            res["groups"][:, 0] = pyoracle.select(gao, flags, res["scores"], params, seed=1)
            res["selection"] = "get_best_record_index replayed in group order, srand(1)"
        keep.update(res)
    return sum(b.n_groups for b in batches), cells, dt, kind


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sp = spec_for(args)
    wl = Workload(dict(sp, filler_bp=0))
    # bounded sample per step so that the whole run stays within a few minutes
    per_step = args.cpu_sample or min(sp["groups"], 128 * threads)
    n_steps = args.warmup + args.steps
    batches = [wl.generate(i * per_step, per_step) for i in range(min(n_steps, 3))]
    times, groups, cells = [], 0, 0
    kind = "port"
    for i in range(n_steps):
        try:
            g, c, dt, kind = cpu_reference_run(wl, [batches[i % len(batches)]], sp["params"], threads)
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
        "config": {"workload": workload_name(sp), "sample": f"{per_step} read groups per step on the host CPU"},
        "gcups": cells / total / 1e9,
        "cpu_baseline": {"value": val, "unit": "read-groups/s", "cores": threads, "kind": kind,
                         "sample": f"{per_step} read groups x {args.steps} steps, thread pool over read groups",
                         "gcups": cells / total / 1e9},
        "e2e": {"value": val, "unit": "read-groups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def parity_at_size(eng, wl, sample, ppreset, threads):
    """The same read groups through the reference's code on the host (thread pool, timed: the CPU baseline of
    this config) and through the GPU path with every debug table: returns (cpu dict, parity dict)."""
    from tools.parity import compare_results
    keep = {}
    g, c, dt, kind = cpu_reference_run(wl, [sample], ppreset, threads, keep=keep)
    cpu = {"value": g / dt, "unit": "read-groups/s", "cores": threads, "kind": kind,
           "sample": f"the first {g} read groups of a step ({dt:.1f} s), thread pool over read groups",
           "gcups": c / dt / 1e9, "seconds": dt}
    eng.rng_seed(1)  # a fresh process: the reference never seeds rand()
    gpu = eng.run_debug(sample, slot=0)
    bad = compare_results(keep, gpu, label="cuda")
    failed = {b.split(":")[0] for b in bad}
    sc, gao = keep["scores"], sample.grp_aln_off
    ties = 0
    for gi in np.flatnonzero(np.diff(gao) > 2):
        sec = sc[gao[gi]:gao[gi + 1]][(sample.flag[gao[gi]:gao[gi + 1]] & 256) != 0]
        ties += int(len(sec) > 1 and (sec == sec.max()).sum() > 1)
    parity = {"read_groups_checked": int(sample.n_groups), "alignments_checked": int(sample.n_alns),
              "scores_bit_exact": "scores(bits)" not in failed,
              "selected_alignment_equal": bool(np.array_equal(gpu["groups"][:, 0], keep["groups"][:, 0])),
              "group_counters_equal": "groups" not in failed or bool(np.array_equal(gpu["groups"][:, 1:], keep["groups"][:, 1:])),
              "baq_at_markers_equal": not ({"markers_baq", "markers_baq_off"} & failed),
              "final_markers_equal": not ({"markers_final", "markers_final_off"} & failed),
              "markers_before_baq_equal": not ({"markers_pre", "markers_pre_off"} & failed),
              "consensus_blocks_equal": not ({"blocks", "block_off"} & failed),
              "extents_equal": "extents" not in failed,
              "hmm_instances_equal": int(gpu["hmm_instances"]) == len(keep["hmm"]),
              "markers_checked": int(len(keep["markers_baq"])), "final_markers_checked": int(len(keep["markers_final"])),
              "groups_with_tied_top_secondaries": ties,
              "selected_secondaries": int((gpu["groups"][:, 0] != gpu["groups"][:, 1]).sum()),
              "selection_oracle": keep.get("selection", "oracle run order"),
              "all_equal": not bad, "mismatches": bad[:4]}
    if bad:
        sys.stderr.write("bench.py: GPU results differ from the CPU reference on the checked batch:\n  " + "\n  ".join(bad) + "\n")
    return cpu, parity


def measure_config(sp, args, steps, warmup, local, rank, world, barrier, allreduce_max, sampler=None, peaks_fp64=None,
                   with_cpu=True):
    """One workload measured through the C ABI: resident arm, end-to-end arm, the HMM launch set alone,
    and (rank 0, N=1) the CPU baseline + at-size parity.  Returns a dict of plain numbers."""
    import torch

    import secphase_b200
    t_setup = time.perf_counter()
    wl = Workload(sp)
    eng = secphase_b200.Secphase(sp["params"], device=local)
    if args.hmm and hasattr(eng, "set_hmm_mode"):
        eng.set_hmm_mode(args.hmm)
    eng.set_reference_codes(wl.codes, wl.off)
    groups, pool = sp["groups"], sp["pool"]
    # shard by query-name range: rank r owns groups [r*span, (r+1)*span)
    span = groups * pool
    # the pools of every batch live in page-locked host memory, as a reader thread decoding straight
    # into sp_host_alloc'ed buffers would leave them (include/secphase_b200.h)
    batches = [secphase_b200.pin_batch(wl.generate(rank * span + i * groups, groups)) for i in range(pool)]
    setup_s = time.perf_counter() - t_setup
    n_slots = min(3, pool)

    # ---- device-resident arm (value) -------------------------------------------------------
    # a slot keeps the batch uploaded to it; with more distinct batches than slots the resident arm cycles
    # the first n_slots of them (the end-to-end arm goes through all of them)
    for s in range(n_slots):
        eng.upload(batches[s], slot=s)

    def resident_steps(n):
        stats, inflight = [], []
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

    resident_steps(warmup)
    barrier()
    t_meas0 = time.perf_counter()
    eng.mark()
    t0 = time.perf_counter()
    stats = resident_steps(steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev_s = device_span_ms(steps) * 1e-3
    dev_wall_s = t1 - t0
    barrier()

    # ---- end-to-end arm (host buffers, H2D + D2H inside the timed region) ------------------
    def e2e_steps(n):
        out, inflight = [], []
        for i in range(n):
            s = i % n_slots
            if len(inflight) == n_slots:
                out.append(eng.wait(inflight.pop(0), copy=True))
            eng.submit(batches[i % len(batches)], slot=s)
            inflight.append(s)
        while inflight:
            out.append(eng.wait(inflight.pop(0), copy=True))
        return out

    # (with more distinct batches than slots every slot meets every batch size: warm up until the slots' device
    # and pinned buffers have grown to the largest, or the timed region would contain their re-allocations)
    e2e_steps(max(warmup, 3) if len(batches) <= n_slots else max(warmup, 3 * len(batches)))
    barrier()
    # the end-to-end arm is a host-visible quantity (pack + H2D + kernels + D2H + result tables in
    # host memory), so it is timed on the host clock around the synchronised region; the device-side
    # span of the same steps is reported next to it
    eng.mark()
    t2 = time.perf_counter()
    estats = e2e_steps(steps)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    e2e_dev_s = device_span_ms(steps) * 1e-3
    barrier()
    e2e_s = t3 - t2

    # ---- the dominant kernel alone: one batch in flight, nothing overlapped, so that the CUDA
    # events around the HMM launches (fork -> class kernels on aux streams -> join, recorded on the
    # slot's stream) time just those kernels
    for s in range(n_slots):  # (the e2e arm left other batches in the slots)
        eng.upload(batches[s], slot=s)
    iso = []
    for i in range(2 + max(3, min(steps, 8))):
        eng.run_resident(0)
        r = eng.wait(0, copy=False)
        if i >= 2:
            iso.append(r)
    clocks = None
    if sampler is not None:
        # the measured phases are short; if the sampler caught fewer than three readings under load,
        # keep the same resident workload running (untimed) until it has
        t_load1 = time.perf_counter()
        while sampler.proc and sampler.count(t_meas0) < 3 and time.perf_counter() - t_load1 < 6.0:
            resident_steps(n_slots)
        clocks = sampler.stop(t_meas0, time.perf_counter())
    if peaks_fp64 is None:  # roofline denominators, measured live
        peaks_fp64 = (eng.fp64_peak(0)[0], eng.fp64_peak(1)[0])
    dfma_ops, dadd_ops = peaks_fp64
    dev_s, e2e_s, dev_wall_s, e2e_dev_s = allreduce_max([dev_s, e2e_s, dev_wall_s, e2e_dev_s])

    groups_total = groups * steps * world
    cells_step = float(np.mean([s["hmm_cells"] for s in stats]))
    hmm_ms = float(np.mean([s["ms_hmm"] for s in iso]))
    iso_cells = float(np.mean([s["hmm_cells"] for s in iso]))
    iso_total_ms = float(np.mean([s["ms_total"] for s in iso]))
    stage_ms = np.mean([s["ms_stage"] for s in iso], axis=0).tolist()
    achieved_tflops = iso_cells * FLOP_PER_CELL / (hmm_ms * 1e-3) / 1e12 if hmm_ms > 0 else 0.0
    peak_tflops = 2.0 * dfma_ops / 1e12
    m = {
        "sp": sp, "steps": steps, "warmup": warmup, "n_slots": n_slots, "setup_s": setup_s,
        "value": groups_total / dev_s, "ms_per_step": 1e3 * dev_s / steps, "dev_wall_ms_per_step": 1e3 * dev_wall_s / steps,
        "gcups_job": cells_step * steps * world / dev_s / 1e9,
        "gcups_kernel": iso_cells / (hmm_ms * 1e-3) / 1e9 if hmm_ms > 0 else 0.0,
        "cells_step": cells_step, "iso_cells": iso_cells, "hmm_ms": hmm_ms, "iso_total_ms": iso_total_ms,
        "hmm_instances_per_step": float(np.mean([s["hmm_instances"] for s in stats])),
        "launches": int(sum(s["gpu_launches"] for s in stats)),
        "e2e_value": groups_total / e2e_s, "e2e_ms_per_step": 1e3 * e2e_s / steps, "e2e_dev_ms_per_step": 1e3 * e2e_dev_s / steps,
        "h2d": int(np.mean([s["h2d_bytes"] for s in estats])), "d2h": int(np.mean([s["d2h_bytes"] for s in estats])),
        "achieved_tflops": achieved_tflops, "peak_tflops": peak_tflops, "dfma_ops": dfma_ops, "dadd_ops": dadd_ops,
        "frac": achieved_tflops / peak_tflops if peak_tflops else None,
        "stage_ms": dict(zip(["h2d", "walk", "group", "emit_sort", "hmm", "score", "d2h"], stage_ms[:7])),
        "sm_partition": dict(zip(("on", "int_sms", "hmm_sms"), eng.sm_partition())),
        "clocks": clocks, "peaks_fp64": peaks_fp64,
        "hmm_mode": eng.hmm_mode() if hasattr(eng, "hmm_mode") else "strict",
        "hmm_rerun_instances_per_step": float(np.mean([s.get("hmm_rerun", 0) for s in iso])),
    }
    if rank == 0 and with_cpu and world == 1:
        threads = os.cpu_count() or 1
        n_par = min(groups, args.parity_groups) if args.parity_groups > 0 else groups
        if sp["title"] == "chr-hifi-30x":
            n_par = groups  # the bench line's own config: the whole step
        try:
            cpu, parity = parity_at_size(eng, wl, batches[0].group_slice(0, n_par), sp["params"], threads)
            m["cpu_baseline"], m["parity"] = cpu, parity
        except RuntimeError as e:
            m["cpu_baseline"], m["parity"] = {"unavailable": str(e)}, None
        if sp.get("tie_parity") and m.get("parity") is not None:
            eng.close()
            tp = sp["tie_parity"]
            wl2 = Workload(dict(sp, synth_over=tp["synth_over"]))
            eng = secphase_b200.Secphase(sp["params"], device=local)
            if args.hmm and hasattr(eng, "set_hmm_mode"):
                eng.set_hmm_mode(args.hmm)
            eng.set_reference_codes(wl2.codes, wl2.off)
            _, ptie = parity_at_size(eng, wl2, wl2.generate(0, tp["groups"]), sp["params"], threads)
            ptie["workload"] = "same generator with " + ", ".join(f"{k}={v:g}" for k, v in tp["synth_over"].items())
            m["parity"]["tie_sample"] = ptie
    eng.close()
    return m


def roofline_dict(m, peaks, main=True):
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    iso_cells, hmm_ms = m["iso_cells"], m["hmm_ms"]
    # Algorithmic HBM bytes of one launch set: the reference window (1 B/base) + the packed query (0.5 B/base)
    # of every instance, SURVEY.md 8(d): ~0.05 B per band cell.  The design's own traffic on top of that
    # (saved forward rows at consumed rows; in strict mode also the per-row scale factors, 8 B out + 8 B back
    # per query row, ~0.35 B/cell) is reported as design_bytes.
    alg_bytes = 0.05 * iso_cells
    design_bytes = (0.4 if m["hmm_mode"] == "strict" else 0.12) * iso_cells
    traffic = traffic_src = None
    for name in ("r02_traffic_k_hmm.json", "r01_traffic_k_hmm2.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tr.get("hmm_mode", "strict") != m["hmm_mode"] or tr.get("workload", "hifi") != m["sp"]["synth"]:
                continue
            traffic = tr["dram_bytes_total"] * iso_cells / tr["band_cells"]
            traffic_src = tr["source"]
            break
        except Exception:
            continue
    r = {"kernel": "k_hmm* (all band-class launches of one step, forked on aux streams)", "bound": "fp64",
         "achieved": m["achieved_tflops"], "peak": m["peak_tflops"], "unit": "TFLOP/s", "frac": m["frac"],
         "traffic": traffic, "traffic_unit": "bytes per launch set (dram__bytes_read.sum + dram__bytes_write.sum)",
         "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
         "algorithmic_bytes_source": "SURVEY.md 8(d): 0.05 B per band cell (reference window + packed query)",
         "design_bytes": design_bytes,
         "flop_per_cell": FLOP_PER_CELL, "hmm_mode": m["hmm_mode"],
         "peak_source": "measured live: register-resident DFMA kernel, 2 flop/instr (MEASURED_PEAKS.json has no FP64 entry)",
         "issue_slot_frac": iso_cells * FLOP_PER_CELL / (hmm_ms * 1e-3) / m["dadd_ops"] if m["dadd_ops"] and hmm_ms > 0 else None,
         "dadd_dmul_ops_per_s": m["dadd_ops"], "kernel_ms_per_step": hmm_ms,
         "kernel_share_of_step": hmm_ms / m["iso_total_ms"] if m["iso_total_ms"] else None,
         "strict_rerun_instances_per_step": m["hmm_rerun_instances_per_step"],
         "timing": "CUDA events on the slot stream around the HMM launches, one batch in flight",
         "hbm": {"algorithmic_gbs": alg_bytes / (hmm_ms * 1e-3) / 1e9 if hmm_ms > 0 else None, "peak_gbs": hbm_peak,
                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    if not main:
        for k in ("traffic_unit", "traffic_source", "peak_source", "timing", "hbm", "algorithmic_bytes_source", "kernel"):
            r.pop(k, None)
    return r


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

    def allreduce_max(vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    # nvidia-smi needs a second or two before its first sample: start it ahead of the warm-up and keep only
    # the samples that arrive inside the measured phases
    sampler = ClockSampler(local)
    sampler.start()
    sp = spec_for(args)
    m = measure_config(sp, args, args.steps, args.warmup, local, rank, world, barrier, allreduce_max, sampler=sampler,
                       with_cpu=not args.no_cpu_baseline)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = {
        "metric": "read_groups_per_sec", "value": m["value"], "unit": "read-groups/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(sp, args.steps), "groups_per_step_per_gpu": sp["groups"], "slots_in_flight": m["n_slots"],
                   "l2": "per-step input (~%.0f MB) exceeds the 126 MB L2; no explicit flush" % (m["h2d"] / 1e6),
                   "parallelism": f"read groups sharded by qname range over {world} GPU(s), no collective",
                   "sm_partition": m["sm_partition"], "hmm_arithmetic": m["hmm_mode"]},
        "gcups": m["gcups_job"], "gcups_kernel": m["gcups_kernel"],
        "hmm_instances_per_step": m["hmm_instances_per_step"],
        "band_cells_per_step": m["cells_step"],
        "e2e": {"value": m["e2e_value"], "unit": "read-groups/s",
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "ms_per_step": m["e2e_ms_per_step"], "device_ms_per_step": m["e2e_dev_ms_per_step"],
                "timing": "host clock around the synchronised region (host packing and result tables are part of "
                          "the call); device_ms_per_step = CUDA-event span of the same steps"},
        "timing": "CUDA events: mark on an idle device -> end of the last batch on each slot stream, max over slots "
                  "and ranks (host wall clock of the same region: %.3f ms/step)" % m["dev_wall_ms_per_step"],
        "gpu_launches": m["launches"],
        "clocks": m["clocks"],
        "roofline": roofline_dict(m, peaks),
        "stage_ms_isolated": m["stage_ms"],
        "setup_s": m["setup_s"],
    }
    if "cpu_baseline" in m:
        line["cpu_baseline"] = m["cpu_baseline"]
        line["parity_vs_cpu_reference"] = m["parity"]
    # ---- the other BASELINE configs (N=1 default run only) ------------------------------------
    if rank == 0 and world == 1 and args.preset == "hifi" and not args.no_per_config and not args.no_cpu_baseline:
        per = {}
        plan = [("hifi_small", 20, 3), ("ont", 15, 3), ("wg_offsets", 15, 3), ("stress", 9, 3)]  # (steps, warm-up): enough steps that the pipeline fill -- one integer phase with idle HMM SMs -- is a few per cent
        for name, k_steps, k_warm in plan:
            t_c = time.perf_counter()
            try:
                mc = measure_config(dict(SPECS[name]), args, k_steps, k_warm, local, rank, world, barrier, allreduce_max,
                                    peaks_fp64=m["peaks_fp64"])
            except Exception as e:  # one config failing must not cost the bench line
                per[name] = {"error": f"{type(e).__name__}: {e}"}
                continue
            per[name] = {
                "baseline_config": SPECS[name]["config"], "workload": workload_name(mc["sp"], k_steps),
                "steps": k_steps, "warmup": k_warm,
                "value": mc["value"], "unit": "read-groups/s", "ms_per_step": mc["ms_per_step"],
                "e2e": {"value": mc["e2e_value"], "h2d_bytes_per_step": mc["h2d"], "d2h_bytes_per_step": mc["d2h"]},
                "gcups": mc["gcups_job"], "gcups_kernel": mc["gcups_kernel"], "band_cells_per_step": mc["cells_step"],
                "roofline": roofline_dict(mc, peaks, main=False), "stage_ms_isolated": mc["stage_ms"],
                "sm_partition": mc["sm_partition"], "gpu_launches": mc["launches"],
                "cpu_baseline": mc.get("cpu_baseline"), "parity_vs_cpu_reference": mc.get("parity"),
                "wall_s": time.perf_counter() - t_c,
            }
        line["per_config"] = per
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
