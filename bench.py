#!/usr/bin/env python
"""bench.py -- RF trees fit/sec on the 1M x 64 scale-up table (BASELINE.json metric / configs[4]).

    python bench.py --gpus N --steps K --warmup W          own arm (CUDA path, libtst.so)
    python bench.py --impl reference --gpus N --steps K ... CPU arm: sklearn RandomForestClassifier (the
                                                           paper's RF in its canonical implementation;
                                                           the reference repo ships no code) on the host cores

A "step" = one pass of the hot path: this rank's 10 folds x TREES_PER_FOLD trees (Philox bootstrap ->
histogram / Gini split / partition growth -> OOF inference), the vote all-gather, and s(x).  Step s
builds trees t in [0, T * world) of every fold with the RNG ids of the job's s-th batch (gid_base = s * 10,000), so
every step is the same amount of work: exactly 10 * T trees per GPU.  Weak scaling: every rank builds TREES_PER_FOLD
(default 64: 640 trees in flight, ~36 GB of the 180 GB HBM; the per-level host round trips are then < 5 % of a step
even on a slow host) trees per fold per step.

job    : strong scaling -- wall-clock of the WHOLE 10-fold x 1000-tree job (host arrays -> bin -> fit -> s(x) on the
         host) through tst.run_tst at N GPUs.
votes_crc32: CRC of the all-reduced votes of trees [0, 4) of every fold (seed 0) -- identical at every N.

value  : trees/s with the binned table resident in HBM (CUDA events, barrier both sides, max over ranks)
e2e    : trees/s through tst.run_tst on HOST arrays -- pinned X/y H2D, bin-edge selection, binning, folds,
         fit, s(x), and D2H of s / votes / folds inside the timed region.
roofline: the growth tier with the largest share of the step (k_grow_big / k_grow_mid4 / k_grow_small4);
         algorithmic bytes = (15 + m) per row-level visit, row visits counted on the device,
         duration = CUDA events around its launches inside libtst (the launching stream).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "test-set-training_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

N_ROWS, N_FEAT, N_CLASSES, K_FOLDS, N_TREES = 1_000_000, 64, 4, 10, 1000
METRIC, UNIT = "rf_trees_fit_per_sec_1Mx64", "trees/s"


def workload(trees_per_fold):
    return dict(workload=f"scale-up table {N_ROWS}x{N_FEAT} f32, {N_CLASSES} classes, {K_FOLDS}-fold x {N_TREES} trees "
                         f"(BASELINE.json configs[4]); step = {K_FOLDS} folds x {trees_per_fold} trees per GPU, full depth, "
                         f"max_features=sqrt(64)=8, 256 bins",
                rows=N_ROWS, features=N_FEAT, classes=N_CLASSES, folds=K_FOLDS, trees_per_fold_per_step=trees_per_fold,
                l2="working set per step (binned tables 128 MB + ~30 MB per tree in flight) is far larger than the 126 MB L2; no explicit flush")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None  # the timed region; only the samples that arrive inside it are summarised

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _nvml_loop(self):
        import pynvml as N

        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        names = [("hw_slowdown", N.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", N.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksThrottleReasonSwPowerCap)]
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        while not self.stop.is_set():
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            pw = N.nvmlDeviceGetPowerUsage(h) / 1e3
            r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((time.time(), [str(sm), str(mx), str(pw)] + ["Active" if r & bit else "Not Active" for _, bit in names]))
            self.stop.wait(0.25)

    def __enter__(self):
        self.mode = os.environ.get("TST_BENCH_CLOCKS", "smi")  # smi = the recipe's nvidia-smi -lms loop; nvml = same counters in-process; off
        if self.mode == "off":
            return self
        if self.mode == "nvml":
            self.stop = threading.Event()
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if getattr(self, "stop", None) is not None:
            self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.t0 is not None and (ts < self.t0 or (self.t1 is not None and ts > self.t1 + 0.3)):
                continue  # nvidia-smi is started before the warm-up so that its start-up does not fall into the timed region
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dram_traffic_per_launch(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture of the current build's launch structure
    (profiles/r02_final_traffic.json, made by tools/traffic_summary.py; the round-1 capture as a fallback)."""
    for name in ("r02_final_traffic.json", "r01_final_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                e = json.load(f).get(kernel, {})
            if e.get("dram_bytes_per_launch") is not None:
                return e.get("dram_bytes_per_launch"), e.get("captured")
        except Exception:
            pass
    return None, None


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import sklearn_tst
    from tst import synth

    cores = sklearn_tst.usable_cores()
    X, y = synth.scaleup_table(N_ROWS, N_FEAT, N_CLASSES, seed=0)
    fold = (np.arange(N_ROWS) % K_FOLDS).astype(np.uint8)
    train = fold != 0
    Xt = X[fold == 0]
    Xtr, ytr = X[train], y[train]  # the fold's training table is prepared once, outside the timed region
    trees = max(1, cores)  # one tree per core per step: the smallest sample that keeps every core busy
    times = []
    for s in range(args.warmup + args.steps):
        r = sklearn_tst.time_fit_sample(Xtr, ytr, None, trees, seed=s, n_jobs=cores, X_test=Xt)
        if s >= args.warmup:
            times.append(r["seconds"])
    total = float(sum(times))
    value = trees * len(times) / total
    import sklearn

    sample = (f"{trees} trees per step on fold 0's {int(train.sum())} training rows + predict_proba of the {len(Xt)} held-out rows, "
              f"sklearn {sklearn.__version__} RandomForestClassifier(n_jobs={cores}) defaults")
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload(args.trees_per_fold), reference_sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample,
                                  note="the reference repository ships no code; this is the committed sklearn restatement of the paper's RF diagnostic"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- own arm
def cpu_baseline_port(X, y, bins_cm, fold):
    """CPU histogram-RF oracle (same algorithm, same Philox streams) on the host cores: bounded sample.
    -> (cpu_baseline dict, oracle stats, oracle votes, trees per fold of the sample)"""
    from oracle import hist_rf as H

    cores = len(os.sched_getaffinity(0))
    prm = H.make_params(N_FEAT)
    per_fold = max(1, (2 * cores + K_FOLDS - 1) // K_FOLDS)
    t = time.perf_counter()
    votes, st, _ = H.fit_oof(bins_cm, y, fold, K_FOLDS, N_TREES, N_CLASSES, prm, seed=0, t0=0, t1=per_fold, nthreads=cores)
    dt = time.perf_counter() - t
    trees = K_FOLDS * per_fold
    return dict(value=trees / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{trees} trees ({per_fold} per fold) of the same job with oracle/hist_rf.c (OpenMP, {cores} threads), {dt:.1f} s"), st, votes, per_fold


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--trees-per-fold", type=int, default=int(os.environ.get("TST_BENCH_TREES_PER_FOLD", "64")))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-job", action="store_true", help="skip the whole-job (10 x 1000 trees) wall-clock")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3  # timing rules: W >= 3

    import torch
    import torch.distributed as dist

    import tst
    from tst import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- setup (untimed): synthetic table, resident binned copy, folds
    T = args.trees_per_fold
    X, y = synth.scaleup_table(N_ROWS, N_FEAT, N_CLASSES, seed=0)
    eng = tst.Engine(local_rank, rank, world)
    eng.bin(X)  # bin edges selected on the device
    classes, yi = tst.api.encode_labels(y)
    y_dev = eng.to_dev(yi, torch.uint8)
    active = torch.ones(N_ROWS, dtype=torch.uint8, device=eng.dev)
    fold_dev = eng.assign_folds(y_dev, active, K_FOLDS, N_CLASSES)
    prm = tst.make_params(N_FEAT)
    votes = eng.new_votes(N_CLASSES)
    m = prm.max_features

    if T * world > N_TREES:
        raise SystemExit(f"bench.py: --trees-per-fold {T} x {world} ranks exceeds the job's {N_TREES} trees per fold")
    BATCH_IDS = K_FOLDS * N_TREES  # RNG ids of one batch of the job: step i uses [i * BATCH_IDS, (i + 1) * BATCH_IDS)

    def step(i):
        # trees t in [0, T*world) of every fold with the RNG ids of batch i; this rank builds exactly T per fold
        # (lid = fold * 1000 + t, lid % world == rank; 1000 is a multiple of 1, 2, 4 and 8)
        eng.fit_oof(y_dev, fold_dev, K_FOLDS, N_TREES, N_CLASSES, prm, seed=0, gid_base=i * BATCH_IDS, t0=0, t1=T * world, votes=votes)
        total = eng.allreduce_votes(votes)
        return eng.finalize(total, y_dev, fold_dev, K_FOLDS, N_TREES, N_CLASSES)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        for i in range(args.warmup):
            step(i)
        barrier()
        eng.stats(reset=True)
        barrier()
        clk.mark_start()
        ev0.record()
        for i in range(args.steps):
            s_dev, _, _ = step(args.warmup + i)
        ev1.record()
        barrier()
        clk.mark_end()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    st = eng.stats()
    trees_per_rank = int(st["trees"])
    trees_total = trees_per_rank * world  # weak scaling: every rank builds the same count
    assert trees_per_rank == K_FOLDS * T * args.steps, (trees_per_rank, K_FOLDS * T * args.steps)
    value = trees_total / (ms / 1e3)
    clocks = clk.summary()

    # ---- world-invariant checksum: trees [0, 4) of every fold, all-reduced; the same bytes at N = 1, 2, 4, 8
    import zlib
    cv = eng.new_votes(N_CLASSES)
    eng.fit_oof(y_dev, fold_dev, K_FOLDS, N_TREES, N_CLASSES, prm, seed=0, t0=0, t1=4, votes=cv)
    cv = eng.allreduce_votes(cv)
    votes_crc = zlib.crc32(cv.cpu().numpy().tobytes())

    # ---- roofline of the dominant kernel: the three growth tiers share the (15 + m) bytes-per-visit accounting of
    # SURVEY 8(d); the one with the largest share of the step is reported, the other two ride along in `tiers`.
    peak, peak_src = measured_hbm_peak()
    tiers = {}
    for name, vis, tms, nl in (("k_grow_big", st["row_visits_big"], st["ms_grow_big"], st["launches_grow_big"]),
                               ("k_grow_wide", st["row_visits_wide"], st["ms_grow_wide"], st["launches_grow_wide"]),
                               ("k_grow_mid4", st["row_visits_mid"], st["ms_grow_mid"], st["launches_grow_mid"]),
                               ("k_grow_small4", st["row_visits_small"], st["ms_grow_small"], st["launches_grow_small"])):
        ab = (15 + m) * vis
        ach = ab / (tms / 1e3) / 1e9 if tms > 0 else 0.0
        tiers[name] = dict(achieved=ach, frac=ach / peak, row_visits=int(vis), ms=tms, launches=int(nl), share_of_step=tms / ms,
                           algorithmic_bytes_per_launch=ab / max(1, nl), avg_launch_ms=tms / max(1, nl))
    dom = max(tiers, key=lambda k: tiers[k]["ms"])
    T_ = tiers[dom]
    traffic, traffic_of = dram_traffic_per_launch(dom)
    roofline = dict(bound="hbm", kernel=dom, achieved=T_["achieved"], peak=peak, unit="GB/s", frac=T_["frac"],
                    traffic=traffic, traffic_launches=traffic_of, peak_source=peak_src,
                    algorithmic_bytes_per_launch=T_["algorithmic_bytes_per_launch"], avg_launch_ms=T_["avg_launch_ms"],
                    launches=T_["launches"], kernel_share_of_step=T_["share_of_step"], tiers=tiers,
                    whole_tree=dict(algorithmic_bytes_per_tree=((15 + m) * (st["row_visits_big"] + st["row_visits_wide"] + st["row_visits_mid"] + st["row_visits_small"])
                                                                + st["inbag_rows"] + 9 * st["oof_steps"]) / max(1, trees_per_rank),
                                    achieved=((15 + m) * (st["row_visits_big"] + st["row_visits_wide"] + st["row_visits_mid"] + st["row_visits_small"]) + st["inbag_rows"]
                                              + 9 * st["oof_steps"]) / (ms / 1e3) / 1e9,
                                    ms_infer=st["ms_infer"], ms_prep=st["ms_prep"], ms_tiny=st.get("ms_grow_tiny", 0.0)),
                    note="level kernels are bound by L2->SM gather latency / issue slots, not by DRAM: the binned table (64 MB) is L2-resident; see DESIGN.md section 4")
    gpu_launches = int(st["kernel_launches"])

    # ---- e2e through the public API on host buffers
    e2e = None
    job = None
    if not args.no_e2e:
        Xp = torch.from_numpy(X).pin_memory()
        yp = y
        h2d = X.nbytes + yi.nbytes
        d2h = N_ROWS * 4 + N_ROWS * N_CLASSES * 4 + N_ROWS
        e2e_steps = max(1, min(args.steps, 3))
        tst.run_tst(Xp, yp, k=K_FOLDS, n_trees=N_TREES, seed=0, tree_range=(0, T * world), engine=eng)  # warm
        barrier()
        t0 = time.perf_counter()
        e2e_trees = 0
        for i in range(e2e_steps):
            res = tst.run_tst(Xp, yp, k=K_FOLDS, n_trees=N_TREES, seed=i + 1, tree_range=(0, T * world), engine=eng)
            e2e_trees += int(res.stats["trees"]) * world  # counted by the library, not assumed
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = dict(value=e2e_trees / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   steps=e2e_steps, ms_per_step=1e3 * dt / e2e_steps, trees=e2e_trees,
                   note="tst.run_tst on host arrays: pinned H2D of X/y, bin-edge selection + binning on the GPU, folds, fit, s(x), D2H of s/votes/folds")
        # ---- strong scaling: the whole 10-fold x 1000-tree job, host arrays in -> s(x) on the host
        if not args.no_job:
            barrier()
            t0 = time.perf_counter()
            res = tst.run_tst(Xp, yp, k=K_FOLDS, n_trees=N_TREES, seed=0, engine=eng)
            barrier()
            dtj = max_over_ranks(time.perf_counter() - t0)
            job = dict(wallclock_s=dtj, trees=K_FOLDS * N_TREES, trees_per_s=K_FOLDS * N_TREES / dtj, scaling="strong",
                       trees_this_rank=int(res.stats["trees"]), blind_accuracy=float(np.nanmean(res.fold_accuracy)),
                       s_crc32=zlib.crc32(res.s.tobytes()),
                       note="whole BASELINE configs[4] job through tst.run_tst: H2D, binning, folds, 10,000 trees sharded over the ranks, vote all-gather, s(x), D2H")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8", data="synthetic", config=workload(T),
                roofline=roofline, e2e=e2e, gpu_launches=gpu_launches, clocks=clocks, job=job, votes_crc32=votes_crc,
                trees_per_step=trees_total // max(1, args.steps), nodes_per_tree=st["nodes"] / max(1, trees_per_rank),
                max_depth=int(st["max_depth"]))

    if rank == 0 and world == 1:
        bins_cm, fold_h = (eng.get_bins(), fold_dev.cpu().numpy()) if not args.no_cpu else (None, None)
        per_fold = max(1, (2 * len(os.sched_getaffinity(0)) + K_FOLDS - 1) // K_FOLDS)  # the CPU sample's trees per fold
        if not args.no_cpu:
            # GPU side of the headline-table parity check, taken while the 1M x 64 table is still the engine's table
            gv = eng.new_votes(N_CLASSES)
            eng.stats(reset=True)
            eng.fit_oof(y_dev, fold_dev, K_FOLDS, N_TREES, N_CLASSES, prm, seed=0, t0=0, t1=per_fold, votes=gv)
            gst = eng.stats()
            gv_h = gv.cpu().numpy().view(np.uint32)
        if not args.no_secondary:  # before the CPU leg: idle OpenMP workers spin for a while and disturb millisecond latencies
            line["secondary"] = secondary_wallclock(tst, synth, eng)
        if not args.no_cpu and not args.no_secondary:
            line["secondary"]["cpu"] = secondary_cpu(synth)
        if not args.no_cpu:
            line["cpu_baseline"], ost, ovotes, per_fold_o = cpu_baseline_port(X, yi, bins_cm, fold_h)
            assert per_fold_o == per_fold
            # parity on the headline table itself: the GPU fit of the sample's trees against the oracle's votes and counters
            same_votes = bool(np.array_equal(gv_h, ovotes))
            same_stats = all(int(gst[k]) == int(ost[k]) for k in ("nodes", "leaves", "searched", "fallback", "inbag_rows", "trees", "oof_steps", "max_depth")) \
                and int(gst["row_visits_big"] + gst["row_visits_wide"] + gst["row_visits_mid"] + gst["row_visits_small"] + gst["row_visits_sub"]) == int(ost["row_visits"])
            line["parity_1Mx64"] = same_votes and same_stats
            line["parity_1Mx64_detail"] = dict(trees=K_FOLDS * per_fold, votes_equal=same_votes, counters_equal=same_stats,
                                               nodes=int(ost["nodes"]), row_visits=int(ost["row_visits"]))
            if per_fold == 4:
                line["votes_crc32_oracle"] = zlib.crc32(ovotes.tobytes())
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def secondary_wallclock(tst, synth, eng):
    """BASELINE.json's second metric: end-to-end TsT + IBP wall-clock on the VSI-shaped table (latency-bound).
    Best of 3 after one warm-up call each (the first call pays allocations and lazy kernel loading)."""
    import torch

    X, y = synth.vsi_table()
    out = {}
    tst.run_tst(X, y, k=5, n_trees=100, seed=0, engine=eng)
    best = 1e9
    for i in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = tst.run_tst(X, y, k=5, n_trees=100, seed=1 + i, engine=eng)
        best = min(best, time.perf_counter() - t)
    out["tst_vsi_5k_x32_5fold_x100_ms"] = 1e3 * best
    out["tst_vsi_kernel_launches"] = int(r.stats["kernel_launches"])
    tst.ibp(X, y, rounds=2, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng)
    best = 1e9
    for i in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        rb = tst.ibp(X, y, rounds=10, k=5, n_trees=200, tau=0.5, batch=250, seed=0, want_s=False, engine=eng)
        best = min(best, time.perf_counter() - t)
    out["ibp_10rounds_5fold_x200_ms"] = 1e3 * best
    out["ibp_removed_per_round"] = rb.removed_per_round.tolist()
    # numeric-answer question type, same table shape: the continuous squared-error criterion ((weight, sum y) histograms on a
    # 16-bit target, packed kernels) and, for comparison, round 1's 32-level criterion on the generic many-class kernels
    Xn, yn = synth.numeric_table(5000, 32, seed=0)
    for crit, tag in (("squared_error", "regression"), ("squared_error_levels", "regression_32levels")):
        tst.run_tst_regression(Xn, yn, k=5, n_trees=100, seed=0, engine=eng, criterion=crit)
        best = 1e9
        for i in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            rr = tst.run_tst_regression(Xn, yn, k=5, n_trees=100, seed=1 + i, engine=eng, criterion=crit)
            best = min(best, time.perf_counter() - t)
        out[f"{tag}_5k_x32_5fold_x100_ms"] = 1e3 * best
        ok = np.isfinite(rr.y_pred)
        out[f"{tag}_oof_r2"] = float(1.0 - ((rr.y_pred[ok] - yn[ok]) ** 2).sum() / ((yn[ok] - yn[ok].mean()) ** 2).sum())
    # per-question-type batched forests (BASELINE.json configs[2]): one forest set per group, all groups in one call
    for kind in ("cvbench", "mmmu", "videomme"):
        Xg, yg, gg, ng = synth.grouped_table(kind)
        tst.run_tst(Xg, yg, k=5, n_trees=100, seed=0, group=gg, engine=eng)
        best = 1e9
        for i in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            rg = tst.run_tst(Xg, yg, k=5, n_trees=100, seed=1 + i, group=gg, engine=eng)
            best = min(best, time.perf_counter() - t)
        out[f"grouped_{kind}_{ng}groups_5fold_x100_ms"] = 1e3 * best
        out[f"grouped_{kind}_trees"] = int(rg.stats["trees"])
    out["note"] = "host arrays in, host arrays out; latency-bound (5k rows): reported as wall-clock, not bandwidth"
    return out


def secondary_cpu(synth):
    """CPU timings beside the secondary metric (BASELINE.md CPU-baseline plan 2-3): sklearn and the C port of the same
    histogram RF on the box's host cores, VSI-shaped TsT (full 5 x 100) and IBP (port: full 10 rounds; sklearn: 2 rounds,
    extrapolated x5 and said so)."""
    from oracle import hist_rf as H
    from oracle import sklearn_tst

    cores = len(os.sched_getaffinity(0))
    X, y = synth.vsi_table()
    out = dict(cores=cores)
    r = sklearn_tst.run_tst(X, y, k=5, n_trees=100, seed=0, n_jobs=cores)
    out["sklearn_tst_vsi_5fold_x100_ms"] = 1e3 * r["seconds"]
    out["sklearn_version"] = r["sklearn_version"]
    t = time.perf_counter()
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
    t_bin = time.perf_counter() - t
    t = time.perf_counter()
    votes, _, _ = H.fit_oof(bins, y, fold, 5, 100, 4, H.make_params(X.shape[1]), seed=0, nthreads=cores)
    H.finalize(votes, y, fold, 5, 100)
    out["port_tst_vsi_5fold_x100_ms"] = 1e3 * (time.perf_counter() - t)
    out["port_binning_ms"] = 1e3 * t_bin
    t = time.perf_counter()
    H.ibp(bins, y, 5, 200, 4, H.make_params(X.shape[1]), seed=0, rounds=10, tau=0.5, batch=250, nthreads=cores, want_s=False)
    out["port_ibp_10rounds_5fold_x200_ms"] = 1e3 * (time.perf_counter() - t)
    r2 = sklearn_tst.ibp(X, y, rounds=2, k=5, n_trees=200, tau=0.5, batch=250, seed=0, n_jobs=cores)
    out["sklearn_ibp_2rounds_5fold_x200_ms"] = 1e3 * r2["seconds"]
    out["sklearn_ibp_10rounds_extrapolated_ms"] = 5e3 * r2["seconds"]
    # numeric answers: sklearn's RandomForestRegressor, 5-fold x 100 trees on the 5k x 32 numeric table (out-of-fold R^2 beside the GPU's)
    from sklearn.ensemble import RandomForestRegressor

    Xn, yn = synth.numeric_table(5000, 32, seed=0)
    foldn = np.arange(len(yn)) % 5
    t = time.perf_counter()
    pn = np.zeros(len(yn))
    for f in range(5):
        pn[foldn == f] = RandomForestRegressor(n_estimators=100, random_state=f, n_jobs=cores).fit(Xn[foldn != f], yn[foldn != f]).predict(Xn[foldn == f])
    out["sklearn_regression_5k_x32_5fold_x100_ms"] = 1e3 * (time.perf_counter() - t)
    out["sklearn_regression_oof_r2"] = float(1.0 - ((pn - yn) ** 2).sum() / ((yn - yn.mean()) ** 2).sum())
    out["note"] = f"host cores = {cores}; sklearn IBP is a 2-round sample extrapolated x5, everything else is the full workload"
    return out


if __name__ == "__main__":
    sys.exit(main())
