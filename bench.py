#!/usr/bin/env python
"""bench.py -- throughput of the MCMC hot path (L-sampler + M-sampler + LOD scoring) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the synthetic
200-member consanguineous pedigree with 10 000 SNPs (swiftlink_b200/synth.py, seed 20261017),
scored at 5 positions per interval every 100th iteration.

One MCMC iteration = MarkovChain::run's loop body (markov_chain.cc:330-361) with the reference's
default sampler mix: with probability 0.5 one L-sampler sweep over all M loci (both parity classes),
otherwise one M-sampler sweep over every meiosis (shuffled; 281 whole-chromosome updates here); every
`scoring period` (100) iterations the descent graph is LOD-scored at all (M-1)*5 positions.
One bench STEP = 100 iterations (the kind of each drawn from the chain's Philox stream, as
GPUMarkovChain::run does) + 1 scoring pass.  metric = MCMC iterations / second, whole job.

With N > 1 GPUs every rank runs an independent replicate chain (the reference's -R, one chain
group per GPU, weak scaling); the LOD accumulators are merged once, at the end of the timed region and
inside it, by small NCCL all-reduces (log-sum-exp: MAX, then SUM of exp(s - max), and SUM of the
counts).  `--config c4` runs BASELINE.json configs[3] instead: a FIXED job of 8 replicates x MC3 ladders
dealt out over the GPUs (strong scaling); `--config east|loop|xlinked` the reference's examples.

--impl reference times the UNMODIFIED reference (oracle/_ref/libswiftref.so, built by
oracle/Makefile) on the host cores with all OpenMP threads: its own LocusSampler sweeps,
MeiosisSampler sweeps and Peeler::process passes (markov_chain.cc:209-266, :342-349, :375-383),
in the loop of MarkovChain::run (mix draw, L- or M-sweep, scoring) on the FULL workload (10 000
SNPs, about a minute of set-up, 7 GB), each step a bounded sample of the bench step (fewer
iterations, the same share of scoring work).  The `cpu_baseline` object of our own line is the
same loop on the first 1000 SNPs (value / 10, stated in its `sample`), so that the default run
stays short.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_MEMBERS, N_MARKERS, N_LOD, SCORING_PERIOD = 200, 10000, 5, 100
LSAMPLER_PROB = 0.5                       # defaults.h: the reference's default -l
# profiling aid only (profiles/*.md say when it was used): a shorter step so that a whole bench run fits under ncu
if os.environ.get("SLK_BENCH_SCORING_PERIOD"):
    SCORING_PERIOD = int(os.environ["SLK_BENCH_SCORING_PERIOD"])
REF_SAMPLE_MARKERS = 1000
METRIC = "mcmc_iterations_per_sec_incl_lod_scoring"
UNIT = "iterations/s"
WORKLOAD = ("synthetic 200-member consanguineous pedigree, 10k SNPs, default sampler mix (-l 0.5), LOD scoring "
            "(5 positions/interval) every %dth iteration" % SCORING_PERIOD)
STEP_DESC = "%d iterations (each an L-sweep w.p. 0.5, else an M-sweep over all meioses) + 1 LOD scoring pass" % SCORING_PERIOD


class quiet_stdout(object):
    """the host classes print what the reference prints (starting likelihood, P(T), ...) on file descriptor 1; the bench
    line must be the only thing on stdout"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved); os.close(self.null)


def init_nccl(local_rank):
    """process group + first collective with file descriptor 1 muted: NCCL prints its version banner there when the
    communicator is created, and the bench line must be the only thing on stdout"""
    import torch
    import torch.distributed as dist
    with quiet_stdout():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        t = torch.zeros(1, device="cuda")
        dist.all_reduce(t)
        torch.cuda.synchronize()
    return dist


def load_order():
    with open(os.path.join(ROOT, "swiftlink_b200", "data", "synth200_peel_order.json")) as f:
        return json.load(f)


def workload_files(n_markers, tag):
    from swiftlink_b200 import synth
    ped = synth.generate(N_MEMBERS, N_MARKERS)
    if n_markers != N_MARKERS:
        ped = synth.subset_markers(ped, n_markers)
    d = os.path.join(tempfile.gettempdir(), "slk_bench_%s_%d" % (tag, os.getpid()))
    return synth.write_linkage(ped, os.path.join(d, "synth"))


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------------

def reference_open(n_markers):
    """the compiled reference on the bench pedigree (first n_markers SNPs), the committed elimination order, one
    sequential-imputation run as start state (the reference's own); returns (Ref, threads, host cores, set-up seconds)"""
    from oracle import refapi
    if not refapi.available():
        return None, 0, 0, 0.0
    t0 = time.time()
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    refapi.set_threads(threads)
    refapi.seed(20261017)
    paths = workload_files(n_markers, "ref")
    r = refapi.Ref(*paths, lodscores=N_LOD)
    names = r.person_names()
    index = dict((n, i) for i, n in enumerate(names))
    order = [index[str(n)] for n in load_order()["order"]]
    assert r.set_peel(order), "reference rejected the committed elimination order"
    r.sequential_imputation(1)
    return r, threads, cores, time.time() - t0


def reference_chain_steps(r, n_markers, steps, warmup, target_step_s):
    """Times `warmup + steps` steps of MarkovChain::run's loop (markov_chain.cc:335-390: mix draw, L- or M-sweep,
    scoring after every SCORING_PERIOD iterations) over the reference's own samplers and peelers.  A step is a BOUNDED
    sample of the bench step (SCORING_PERIOD iterations + 1 scoring pass): `it` iterations and a scoring pass over
    it / SCORING_PERIOD of the intervals, i.e. the same work per iteration, with `it` sized from the first step's pace.
    Returns (iterations/s over the timed steps, info)."""
    # pace: one iteration of each kind, one scheduler trial of every kind (optimal_num_lgroups, markov_chain.cc:269-311)
    t, lgroups, counts = r.bench_chain(2, 2, LSAMPLER_PROB, max(1, (n_markers - 1) * 2 // SCORING_PERIOD), 1)
    it_step = int(min(SCORING_PERIOD, max(2, round(target_step_s / max(t / 2.0, 1e-6)))))
    lod_int = max(1, int(round((n_markers - 1) * it_step / float(SCORING_PERIOD))))
    times, kinds = [], [0, 0, 0]
    for s in range(warmup + steps):
        t, _, counts = r.bench_chain(it_step, it_step, LSAMPLER_PROB, lod_int, 0 if lgroups < 0 else -lgroups)
        if s >= warmup:
            times.append(t)
            for k in range(3):
                kinds[k] += counts[k]
    secs = float(np.sum(times))
    info = dict(iterations_per_step=it_step, lod_intervals_per_step=lod_int, lgroups=lgroups, seconds=secs,
                l_sweeps=kinds[0], m_sweeps=kinds[1], scoring_passes=kinds[2], ms_per_step=1e3 * secs / max(steps, 1))
    return steps * it_step / secs, info


def reference_measure(steps, warmup, n_markers=REF_SAMPLE_MARKERS, target_step_s=5.0):
    """Times the reference's own CPU path on the first n_markers SNPs of the bench pedigree.  Returns
    (iterations/s at that marker count, info); the caller states the sample."""
    r, threads, cores, t_setup = reference_open(n_markers)
    if r is None:
        return None, dict(error="oracle/_ref/libswiftref.so not built")
    value, info = reference_chain_steps(r, n_markers, steps, warmup, target_step_s)
    # the three components alone, for the per-kernel comparisons (trait positions/s, locus updates/s)
    t_old = r.bench_lsweeps(1, -1)
    t_grp = r.bench_lsweeps(1, 4) if threads > 1 else t_old
    t_sweep = min(t_old, t_grp)
    t_lod = r.bench_lodpasses(1) if n_markers <= 2000 else None       # a full pass at 10k SNPs takes tens of seconds
    t_ms = r.bench_msweeps(1)
    info.update(cores=threads, host_cores=cores, n_markers=n_markers, setup_s=t_setup, l_sweep_s=t_sweep, m_sweep_s=t_ms,
                lod_pass_s=t_lod, locus_updates_per_s=n_markers / t_sweep,
                trait_positions_per_s=((n_markers - 1) * N_LOD / t_lod) if t_lod else None)
    r.close()
    return value, info


def run_reference(args, rank):
    if rank != 0:
        return
    t0 = time.time()
    # the whole run (set-up about 40 s + warm-up and timed steps) has to end within a few minutes whatever --steps is
    step_s = min(args.ref_step_seconds, 120.0 / max(args.steps + args.warmup, 1))
    value, info = reference_measure(args.steps, args.warmup, N_MARKERS, target_step_s=step_s)
    if value is None:
        print(json.dumps({"impl": "reference", "unavailable": info["error"]}))
        return
    sample = ("UNMODIFIED reference (oracle/_ref/libswiftref.so) on the full workload (%d members x %d SNPs), %d OpenMP threads: "
              "the loop of MarkovChain::run over its own LocusSampler / MeiosisSampler / Peeler objects; one step = %d "
              "iterations (L-sweep w.p. 0.5 with the scheduler its own timing picks, else M-sweep) + a scoring pass over "
              "%d of the %d intervals, i.e. the bench step's work per iteration (scoring every %dth) on a bounded sample; "
              "timed steps held %d L-sweeps, %d M-sweeps, %d scoring passes; set-up %.0f s not timed" %
              (N_MEMBERS, N_MARKERS, info["cores"], info["iterations_per_step"], info["lod_intervals_per_step"],
               N_MARKERS - 1, SCORING_PERIOD, info["l_sweeps"], info["m_sweeps"], info["scoring_passes"], info["setup_s"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "detail": {k: info[k] for k in ("iterations_per_step", "lod_intervals_per_step", "lgroups", "l_sweeps", "m_sweeps",
                                        "scoring_passes", "l_sweep_s", "m_sweep_s", "locus_updates_per_s", "host_cores",
                                        "setup_s")},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def bench_config():
    """the `config` object both arms print"""
    return {"workload": WORKLOAD, "step": STEP_DESC, "lsampler_prob": LSAMPLER_PROB,
            "n_members": N_MEMBERS, "n_markers": N_MARKERS, "n_lod": N_LOD}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read)
            self.thread.daemon = True
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args, rank, world, local_rank):
    import torch
    from swiftlink_b200 import build, capi, host as H

    build.build()
    if not torch.cuda.is_available() or capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        dist = init_nccl(local_rank)

    t_setup = time.time()
    paths = workload_files(N_MARKERS, "r%d" % rank)
    hst = H.Host(*paths, lodscores=N_LOD)
    assert hst.set_peel_by_names(load_order()["order"]), "committed elimination order rejected"
    plan = H.PlanFromHost(hst, device=local_rank)
    stats = plan.stats()
    chain = capi.Chain(plan, seed=20261017, chain_id=rank)     # replicate r on GPU r
    # a dedicated (non-default) stream shared by torch's events and the library's launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    chain.set_stream(stream.cuda_stream)
    M, N = hst.M, hst.N

    # start state: one sequential-imputation run (LocusSampler::start_from), then a short burn-in
    si_weight = chain.sequential_imputation(run=0, start_locus=M // 2)
    it = 1
    for _ in range(10):
        chain.lsampler_sweep(it); it += 1
    chain.msampler_sweep(it); it += 1
    chain.sync()
    trait_prob = plan.trait_likelihood()
    t_setup = time.time() - t_setup

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    pinned_dg = torch.zeros((M, N, 2), dtype=torch.int32).pin_memory()
    pinned_lod = torch.zeros(((M - 1) * N_LOD,), dtype=torch.float64).pin_memory()
    chain.dg_download_ptr(pinned_dg.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    n_meioses = len(plan.msampler_ordering())
    launches = [0]
    kinds = [0, 0]
    from swiftlink_b200 import dist as sdist
    merged, merge_s = [None], [0.0]

    def hot_step(start_it):
        for k in range(SCORING_PERIOD):
            if chain.sweep_is_lsampler(start_it + k, LSAMPLER_PROB):
                chain.lsampler_sweep(start_it + k)
                launches[0] += 2
                kinds[0] += 1
            else:
                chain.msampler_sweep(start_it + k)
                launches[0] += 1 + 2 * ((n_meioses + 1) // 2)
                kinds[1] += 1
        chain.lodscore_accumulate()
        launches[0] += 1
        return start_it + SCORING_PERIOD

    def merge_now():
        """N > 1: the replicates' tables merged as LODscores::merge_results does (MAX / SUM-of-exp / SUM all-reduces over
        NCCL), once per timed region and inside it, as the reference merges once per job (linkage_program.cc:96-108)"""
        if dist is None:
            return
        chain.sync()                                       # (so that merge_ms is the merge, not the last step's tail)
        t1 = time.perf_counter()
        cnt = chain.lodscore_read_ptr(pinned_lod.data_ptr())
        merged[0] = sdist.merge_lod(pinned_lod.to("cuda", non_blocking=True), cnt)
        torch.cuda.synchronize()
        merge_s[0] += time.perf_counter() - t1

    def timed(step_fn, k_steps, it0):
        """EXACTLY K steps back to back inside one bracket (barrier + synchronize on both sides), CUDA events on the
        launching stream; the merge of the replicates' tables (N > 1) is inside the bracket.  No flush between the
        steps: a step touches far more than the L2 holds (388 MB of peel scratch per L-sampler launch alone), and one
        256 MiB flush precedes the bracket.  Returns (seconds, next iteration)."""
        flush.fill_(1)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k_steps):
            it0 = step_fn(it0)
        merge_now()
        b.record(stream)
        b.synchronize()
        barrier()
        return a.elapsed_time(b) * 1e-3, it0

    # ---- device-resident throughput ------------------------------------------------------------
    chain.lodscore_init()
    for _ in range(args.warmup):
        it = hot_step(it)
    chain.sync()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches[0] = 0
    kinds[0] = kinds[1] = 0
    merge_s[0] = 0.0
    secs, it = timed(hot_step, args.steps, it)
    merge_ms_per_step = 1e3 * merge_s[0]
    timed_launches, timed_kinds = launches[0], list(kinds)
    chain.sync()
    clock_info = clocks.stop()
    if dist is not None:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    value = world * args.steps * SCORING_PERIOD / secs

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    h2d = pinned_dg.numel() * 4
    d2h = pinned_dg.numel() * 4 + pinned_lod.numel() * 8

    def e2e_step(start_it):
        chain.dg_upload_ptr(pinned_dg.data_ptr())             # int32[M][N][2], as GPULodscores::calculate copies it
        nxt = hot_step(start_it)
        chain.dg_download_ptr(pinned_dg.data_ptr())
        chain.lodscore_read_ptr(pinned_lod.data_ptr())
        return nxt

    def timed_wall(step_fn, k_steps, it0):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_steps):
            it0 = step_fn(it0)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        barrier()
        return total, it0

    for _ in range(min(args.warmup, 2)):
        it = e2e_step(it)
    e2e_secs, it = timed_wall(e2e_step, args.steps, it)
    if dist is not None:
        t = torch.tensor([e2e_secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_secs = float(t.item())
    e2e_value = world * args.steps * SCORING_PERIOD / e2e_secs

    # ---- per-kernel durations (CUDA events on the launching stream) -------------------------------------
    def kernel_ms(fn, reps):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b) / reps

    ctr = [it]

    def one_window():
        chain.lsampler_window(ctr[0], 2, ctr[0] & 1)
        ctr[0] += 1
    ls_ms = kernel_ms(one_window, 40)                              # one launch = M/2 loci
    lod_ms = kernel_ms(chain.lodscore_accumulate, 5)               # one launch = (M-1)*5 positions
    mctr = [it + 100000]

    def one_msweep():
        chain.msampler_sweep(mctr[0])
        mctr[0] += 1
    msweep_ms = kernel_ms(one_msweep, 3)                           # 1 + 2 * ceil(n_meioses / 2) launches
    order = plan.msampler_ordering()
    ms_lik_ms = kernel_ms(lambda: chain.debug_msampler_launch(order[0], order[1], 0, 20), 1) / 20.0
    ms_chain_ms = kernel_ms(lambda: chain.debug_msampler_launch(order[2], order[3], 1, 20), 1) / 20.0
    chain.sync()
    fp64_peak = capi.measure_fp64_peak(local_rank)

    # ---- several replicate chains in flight on this GPU (the -R loop of swift, run_replicates in csrc/host/gpu.cc) ----
    # One chain leaves most of the SMs idle during an M-sweep, so replicates advanced in turn on their own streams
    # overlap.  Reported under `derived` only: the headline stays the single chain BASELINE.json's metric is quoted on.
    in_flight = {}
    if args.in_flight > 1:
        K = args.in_flight
        dg_now = chain.dg_download()
        extra = []
        for k in range(1, K):
            c = capi.Chain(plan, seed=20261017, chain_id=1000 * (rank + 1) + k)
            st = torch.cuda.Stream()
            c.set_stream(st.cuda_stream)
            c.dg_upload(dg_now)
            c.lodscore_init()
            extra.append((c, st))
        group = [(chain, stream)] + extra

        def group_step(start_it):
            for k in range(SCORING_PERIOD):
                for c, _ in group:
                    if c.sweep_is_lsampler(start_it + k, LSAMPLER_PROB):
                        c.lsampler_sweep(start_it + k)
                    else:
                        c.msampler_sweep(start_it + k)
            for c, _ in group:
                c.lodscore_accumulate()
            return start_it + SCORING_PERIOD

        it = group_step(it)                                          # warm-up
        torch.cuda.synchronize()
        g_steps, g_ms = 2, 0.0
        for _ in range(g_steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _, st in extra:
                st.wait_event(a)
            it = group_step(it)
            ends = []
            for _, st in group:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ends.append(e)
            torch.cuda.synchronize()
            g_ms += max(a.elapsed_time(e) for e in ends)
        in_flight = {"chains": K, "iterations_per_s_all_chains": K * g_steps * SCORING_PERIOD / (g_ms * 1e-3),
                     "ms_per_step": g_ms / g_steps,
                     "note": "%d replicate chains of the same workload advanced in turn, one stream each, on one GPU; "
                             "aggregate over the chains, CUDA events, max over the streams" % K}
        for c, _ in extra:
            c.close()

    # ---- merge replicates (LODscores::merge_results, lod_score.h:98-105) -----------------------------------
    raw, count = chain.lodscore_read()
    raw = torch.from_numpy(raw.ravel().copy()).cuda()
    raw, merged_count = sdist.merge_lod(raw, count)
    lod = sdist.normalise(raw, merged_count, trait_prob)
    lod_max = float(lod.max().item())
    lod_argmax = int(lod.argmax().item()) // N_LOD

    # ---- roofline --------------------------------------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (IOError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = {}
    for tag in ("r2", "r1"):
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % tag)) as f:
                traffic = json.load(f)
            break
        except (IOError, ValueError):
            pass
    loci_per_launch = M / 2.0
    ls_flops = stats["flops_ls"] * loci_per_launch
    ls_tflops = ls_flops / (ls_ms * 1e-3) / 1e12
    lod_flops = stats["flops_lod"] * (M - 1) * N_LOD
    lod_tflops = lod_flops / (lod_ms * 1e-3) / 1e12
    nf = N - hst.F
    # algorithmic bytes (minimal encodings): per locus update read 2x2(N-F) bits of
    # neighbouring meioses + 2N bits of genotypes + 16 B theta, write 2(N-F) bits
    ls_bytes = loci_per_launch * ((4 * nf + 2 * N + 2 * nf) / 8.0 + 16)
    # The dominant kernel of the default mix is the M-sampler's (incremental) likelihood kernel: integer / byte
    # work, one thread per (locus, hypothesis).  Algorithmic bytes per evaluation (minimal encodings): the graph
    # row 2(N-F) bits, the typed people's genotypes 2 nt bits, two allele-frequency logs 16 B, ln L out 8 B.
    n_typed = int(np.asarray(hst.person_table()["typed"]).sum())
    ms_bytes_per_eval = (2 * nf + 2 * n_typed) / 8.0 + 24.0
    ms_evals = 3 * M                                               # three hypotheses per locus per launch
    ms_bytes = ms_bytes_per_eval * ms_evals
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = 1e6 * float(clock_info.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)

    def issue_bound(kernel, launch_ms):
        """instruction-issue roofline of a kernel: warp instructions per launch (ncu, smsp__inst_executed.sum of the committed
        capture of the same launch shape) / live launch time, against 4 issue slots per SM per clock"""
        n = traffic.get("_warp_inst_per_launch", {}).get(kernel)
        if not n or not launch_ms:
            return None
        peak = 4.0 * sm_count * sm_hz
        return {"bound": "issue slots", "achieved": n / (launch_ms * 1e-3), "peak": peak, "unit": "warp instructions/s",
                "frac": n / (launch_ms * 1e-3) / peak, "warp_instructions_per_launch": n,
                "ncu_issue_active_pct": traffic.get("_issue_active_pct", {}).get(kernel)}
    roofline = {
        "kernel": "slk_ms_step_kernel", "bound": "hbm", "achieved": ms_bytes / (ms_lik_ms * 1e-3) / 1e9, "peak": hbm_peak,
        "unit": "GB/s", "frac": ms_bytes / (ms_lik_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get("slk_ms_step_kernel"),
        "traffic_source": traffic.get("_source"),
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
        "algorithmic_bytes_per_launch": ms_bytes, "launch_ms": ms_lik_ms, "units_per_launch": ms_evals,
        "note": "latency-bound integer walk (one dependent chain per thread, 939 warps per launch): neither HBM nor a math "
                "pipe is the limiter -- the bound that describes it is `issue` below; in a sweep two of these launches are in "
                "flight and overlap the chain kernels (programmatic dependent launch), launch_ms is the kernel timed alone; "
                "see DESIGN.md section 4",
        "issue": issue_bound("slk_ms_step_kernel", ms_lik_ms),
        "ncu": {"issue_slots_busy_pct": traffic.get("_issue_active_pct", {}), "fp64_pipe_busy_pct": traffic.get("_fp64_pipe_pct", {}),
                "source": traffic.get("_issue_source")},
        "ms_chain_kernel": {"launch_ms": ms_chain_ms, "steps_per_launch": 2, "traffic": traffic.get("slk_ms_chain_kernel")},
        "lsampler_kernel": {
            "bound": "fp64", "achieved": ls_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ls_tflops / fp64_peak if fp64_peak else None,
            "peak_source": "FP64 FMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)",
            "algorithmic_flops_per_launch": ls_flops, "launch_ms": ls_ms, "units_per_launch": loci_per_launch,
            "traffic": traffic.get("slk_lsampler_kernel"), "issue": issue_bound("slk_lsampler_kernel", ls_ms),
            "hbm": {"achieved": ls_bytes / (ls_ms * 1e-3) / 1e9, "frac": ls_bytes / (ls_ms * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_bytes_per_launch": ls_bytes}},
        "lodscore_kernel": {"bound": "fp64", "achieved": lod_tflops, "frac": lod_tflops / fp64_peak if fp64_peak else None,
                            "launch_ms": lod_ms, "algorithmic_flops_per_launch": lod_flops, "traffic": traffic.get("slk_lodscore_kernel"),
                            "issue": issue_bound("slk_lodscore_kernel", lod_ms),
                            "trait_positions_per_s": (M - 1) * N_LOD / (lod_ms * 1e-3)},
        "msampler": {"sweep_ms": msweep_ms, "meioses_per_sweep": n_meioses, "us_per_meiosis_step": 1e3 * msweep_ms / max(n_meioses, 1),
                     "locus_likelihoods_per_s": 1.5 * n_meioses * M / (msweep_ms * 1e-3)},
    }
    n_l, n_m = timed_kinds
    t_l, t_m, t_s = n_l * 2 * ls_ms, n_m * msweep_ms, args.steps * lod_ms
    roofline["share_of_step"] = {"slk_lsampler_kernel": t_l / (t_l + t_m + t_s), "slk_ms_kernels": t_m / (t_l + t_m + t_s),
                                 "slk_lodscore_kernel": t_s / (t_l + t_m + t_s)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": STEP_DESC, "lsampler_prob": LSAMPLER_PROB,
                   "l_sweeps_timed": n_l, "m_sweeps_timed": n_m, "meioses_per_m_sweep": n_meioses,
                   "n_members": N, "n_founders": hst.F, "n_markers": M, "n_lod": N_LOD, "parallelism": "replicate chain per GPU" + ("; LOD tables merged over NCCL at the end of the timed region, inside it" if world > 1 else ""),
                   "l2": "inputs larger than L2: a step streams 388 MB of peel scratch per L-sampler launch and 100 MB per scoring pass; one 256 MiB flush before the timed region",
                   "peel_cost_sum4c": stats["sum_cells"], "max_cutset": stats["max_cutset"],
                   "ls_team_threads": stats["ls_team_threads"], "lod_team_threads": stats["lod_team_threads"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_secs / args.steps},
        "gpu_launches": timed_launches,
        "clocks": clock_info,
        "roofline": roofline,
        "derived": {"locus_updates_per_s": world * M * n_l / secs,
                    "trait_positions_per_s_kernel": (M - 1) * N_LOD / (lod_ms * 1e-3),
                    "l_sweep_ms": 2 * ls_ms, "m_sweep_ms": msweep_ms, "lod_pass_ms": lod_ms,
                    "lod_max": lod_max, "lod_argmax_interval": lod_argmax, "scoring_passes_merged": merged_count,
                    "setup_s": t_setup, "si_log10_weight": si_weight / np.log(10.0),
                    "merge_ms_incl_wait_for_slowest_rank": merge_ms_per_step if world > 1 else None,
                    "replicates_in_flight": in_flight},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info = reference_measure(2, 1, REF_SAMPLE_MARKERS, target_step_s=4.0)
            if v is not None:
                scale = float(REF_SAMPLE_MARKERS) / N_MARKERS
                line["cpu_baseline"] = {
                    "value": v * scale, "unit": UNIT, "cores": info["cores"], "kind": "reference",
                    "sample": "UNMODIFIED reference (oracle/_ref/libswiftref.so), %d OpenMP threads, the loop of MarkovChain::run "
                              "(mix draw, L- or M-sweep, scoring every %dth iteration) on the first %d of the %d SNPs of the same "
                              "pedigree: 2 timed steps of %d iterations + scoring of %d intervals; per-locus cost does not depend "
                              "on M (checked at 1k / 2k / 10k SNPs), value = measured / %d; `bench.py --impl reference` runs the "
                              "full 10k-SNP workload" % (info["cores"], SCORING_PERIOD, REF_SAMPLE_MARKERS, N_MARKERS,
                                                         info["iterations_per_step"], info["lod_intervals_per_step"],
                                                         int(round(1.0 / scale))),
                    "measured_at_sample": v,
                    "trait_positions_per_s": info["trait_positions_per_s"],
                    "locus_updates_per_s": info["locus_updates_per_s"],
                    "l_sweep_s_at_sample": info["l_sweep_s"], "m_sweep_s_at_sample": info["m_sweep_s"],
                    "lod_pass_s_at_sample": info["lod_pass_s"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": info["error"]}
        except Exception as e:                                       # the baseline must not take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    chain.close()
    plan.close()
    hst.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



# ---------------------------------------------------------------------------------------------
# the reference's example configurations (BASELINE.json configs[0], [1], [4]): --config east | loop | xlinked
# ---------------------------------------------------------------------------------------------

SMALL_CONFIGS = {
    # name: replicates (-R), -X, iterations per bench step, reference CLI command
    # ref_iterations: burn-in + iterations per replicate of the reference CLI run (its default is 100 000)
    "east":    dict(runs=1, sex_linked=False, step_iterations=2000, ref_iterations=10000, cli="swift -p east.ped -m east.map -d east.dat"),
    "loop":    dict(runs=10, sex_linked=False, step_iterations=2000, ref_iterations=100000, cli="swift -p loop.ped -m loop.map -d loop.dat -R 10"),
    "xlinked": dict(runs=1, sex_linked=True, step_iterations=2000, ref_iterations=100000, cli="swift -p xlinked.ped -m xlinked.map -d xlinked.dat -X"),
}
SMALL_SCORING_PERIOD = 10                  # defaults.h: DEFAULT_MCMC_SCORING_PERIOD
SMALL_SI_RUNS = 100


def small_config_desc(name):
    c = SMALL_CONFIGS[name]
    return {"workload": "reference example %s (`%s`), default sampler mix (-l 0.5), LOD scoring (5 positions/interval) every "
                        "%dth iteration, %d replicate chain(s)" % (name, c["cli"], SMALL_SCORING_PERIOD, c["runs"]),
            "step": "%d MCMC iterations of each of the %d replicate(s), scoring passes included" % (c["step_iterations"], c["runs"]),
            "lsampler_prob": LSAMPLER_PROB, "n_lod": N_LOD, "replicates": c["runs"], "sex_linked": c["sex_linked"]}


def reference_cli_best(name):
    """the reference CLI with one thread and with every host core (`-c`): on the tiny examples its OpenMP loops lose to one
    thread; the faster of the two is the baseline, with the thread count it used"""
    its = SMALL_CONFIGS[name]["ref_iterations"]
    cores = min(os.cpu_count() or 1, 64)
    # (the all-core run is given a fifth of the iterations: on these inputs it is the slower one by far)
    runs = [reference_cli_rate(name, th, its if th == 1 else max(its // 5, 1000)) for th in sorted(set([1, cores]))]
    v, info = max(runs, key=lambda r: r[0])
    info["all_runs"] = [dict(threads=i["threads"], iterations_per_s=r, setup_s=i["setup_s"], run_s=i["run_s"]) for r, i in runs]
    return v, info


def reference_cli_rate(name, threads, iterations):
    """the reference's own `swift` (oracle/_ref/swift, unmodified, built by oracle/Makefile) on the example: wall time of
    the run minus the wall time of the same command with `-b 0 -i 1` (set-up: parsing, peel-sequence search,
    sequential imputation), as BASELINE.md section 3.2 defines the end-to-end rate"""
    from oracle import refapi
    exe = os.path.join(refapi.REF_DIR, "swift")
    c = SMALL_CONFIGS[name]
    ped, mp, dat = refapi.example(name)
    d = tempfile.mkdtemp(prefix="slk_refcli_")
    base = [exe, "-p", ped, "-m", mp, "-d", dat, "-c", str(threads), "-o", os.path.join(d, "out.txt")]
    if c["runs"] > 1:
        base += ["-R", str(c["runs"])]
    if c["sex_linked"]:
        base += ["-X"]

    def wall(burnin, iters):
        t0 = time.perf_counter()
        p = subprocess.run(base + ["-b", str(burnin), "-i", str(iters)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        if p.returncode != 0:
            raise RuntimeError("reference swift failed: " + p.stderr[-300:])
        return time.perf_counter() - t0
    t_setup = wall(0, 1)
    t_run = wall(iterations // 2, iterations - iterations // 2)
    its = c["runs"] * iterations
    return its / max(t_run - t_setup, 1e-9), dict(setup_s=t_setup, run_s=t_run, iterations=its, threads=threads,
                                                  command=" ".join(["swift"] + base[1:7] + base[9:]) + " -b %d -i %d" % (iterations // 2, iterations - iterations // 2))


def run_small_reference(args, rank):
    if rank != 0:
        return
    from oracle import refapi
    name = args.config
    if not (refapi.available() and os.path.exists(os.path.join(refapi.REF_DIR, "swift"))):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/swift not built"}))
        return
    best, info = reference_cli_best(name)
    threads = info["threads"]
    line = {"impl": "reference", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * SMALL_CONFIGS[name]["runs"] * SMALL_CONFIGS[name]["step_iterations"] / best,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "reference example files",
            "config": small_config_desc(name),
            "cpu_baseline": {"value": best, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "UNMODIFIED reference CLI `%s`: wall time minus the wall time of the same command with "
                                       "-b 0 -i 1 (set-up %.2f s)" % (info["command"], info["setup_s"])},
            "e2e": {"value": best, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "detail": info}
    print(json.dumps(line))


def run_small_ours(args, rank, world, local_rank):
    """one of the reference's example pedigrees through the host API the `swift` command line drives
    (swiftlink::run_replicates, csrc/host/gpu.cc): every replicate chain on the device, all of them in flight at once"""
    import torch
    from oracle import refapi                       # example input files only (they live in oracle/_ref/examples)
    from swiftlink_b200 import build, capi, host as H
    build.build()
    if not torch.cuda.is_available() or capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    name = args.config
    cfg = SMALL_CONFIGS[name]
    files = refapi.example(name)
    hst = H.Host(*files, sex_linked=cfg["sex_linked"], lodscores=N_LOD)
    hst.build_peel(100000)
    S, R = cfg["step_iterations"], cfg["runs"]

    def call(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with quiet_stdout():
            lod = hst.run_replicates(R, R, 0, iters, scoring_period=SMALL_SCORING_PERIOD, seed=20261017 + rank, device=local_rank,
                                     lsampler_prob=LSAMPLER_PROB, si_iterations=SMALL_SI_RUNS)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, lod
    for _ in range(max(args.warmup, 1)):
        call(S // 10)
    setups = sorted(call(1)[0] for _ in range(3))
    t_setup = setups[1]
    clocks = ClockSampler(local_rank)
    clocks.start()
    secs, lod = 0.0, None
    for _ in range(args.steps):
        t, lod = call(S)
        secs += max(t - t_setup, 1e-9)
    clock_info = clocks.stop()
    if world > 1:
        dist = init_nccl(local_rank)
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    value = world * R * S * args.steps / secs

    # ---- the LOD kernel alone on one descent graph, next to the reference's own GPU kernel and its CPU peeler ----
    derived = {"setup_s_per_call": t_setup, "lod_max": float(np.max(lod)),
               "note": "value = replicates x iterations / (wall time of swiftlink::run_replicates - wall time of the same call "
                       "with 1 iteration); host buffers in and out, so value and e2e are the same measurement"}
    plan = H.PlanFromHost(hst, device=local_rank)
    chain = capi.Chain(plan, seed=1)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    chain.set_stream(stream.cuda_stream)
    chain.sequential_imputation(run=0, start_locus=hst.M // 2)
    for it in range(1, 20):
        chain.lsampler_sweep(it)
    chain.sync()

    def kernel_ms(fn, reps):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b) / reps
    positions = (hst.M - 1) * N_LOD
    lod_ms = kernel_ms(chain.lodscore_accumulate, 200)
    ctr = [100]

    def lsweep():
        chain.lsampler_sweep(ctr[0]); ctr[0] += 1

    def msweep():
        chain.msampler_sweep(ctr[0]); ctr[0] += 1
    ls_ms = kernel_ms(lsweep, 200)
    ms_ms = kernel_ms(msweep, 50)
    derived.update({"lod_pass_ms": lod_ms, "trait_positions_per_s_kernel": positions / (lod_ms * 1e-3), "l_sweep_ms": ls_ms,
                    "m_sweep_ms": ms_ms, "meioses_per_m_sweep": len(plan.msampler_ordering()),
                    "n_members": hst.N, "n_markers": hst.M})
    stats = plan.stats()
    fp64_peak = capi.measure_fp64_peak(local_rank)
    lod_tflops = stats["flops_lod"] * positions / (lod_ms * 1e-3) / 1e12
    roofline = {"kernel": "slk_lodscore_kernel", "bound": "tensor", "unit": "TFLOP/s", "achieved": lod_tflops, "peak": fp64_peak,
                "frac": lod_tflops / fp64_peak if fp64_peak else None, "traffic": None,
                "peak_source": "FP64 FMA microbenchmark run in this process (the kernel is FP64 CUDA-core work, not tensor-core work; "
                               "MEASURED_PEAKS.json has no FP64 figure)",
                "launch_ms": lod_ms, "units_per_launch": positions,
                "note": "%d positions per launch: a launch of this size is launch-latency bound, not pipe bound" % positions}
    chain.close(); plan.close()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "reference example files", "config": small_config_desc(name),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": R * hst.M * hst.N * 8,
                    "d2h_bytes_per_step": R * hst.M * hst.N * 8 + positions * 8, "ms_per_step": 1e3 * secs / args.steps},
            "gpu_launches": None, "clocks": clock_info, "roofline": roofline, "derived": derived}
    # launches of a step, counted from the sweep kinds the chains draw: L-sweep 2, M-sweep 1 + 2 ceil(n / 2), scoring 1
    n_me = derived["meioses_per_m_sweep"]
    line["gpu_launches"] = int(args.steps * R * (S * (0.5 * 2 + 0.5 * (1 + 2 * ((n_me + 1) // 2))) + S // SMALL_SCORING_PERIOD))
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import refgpu
            if refgpu.available() and not cfg["sex_linked"]:         # the reference refuses -g with -X (main.cc:534-537)
                out = os.path.join(tempfile.gettempdir(), "slk_refgpu_%s_%d.npz" % (name, os.getpid()))
                rg = refgpu.run_in_subprocess(name, 50, out)
                derived["reference_gpu_kernel"] = {
                    "kernel": "lodscore_kernel (cuda_lodscore.cu:389-467) via GPULodscores::calculate, compiled for sm_100a",
                    "ms_per_pass": 1e3 * float(rg["secs_per_pass"]), "trait_positions_per_s": positions / float(rg["secs_per_pass"]),
                    "block_threads": int(rg["block_threads"]), "setup_s": float(rg["setup_s"]),
                    "reference_cpu_1_thread_trait_positions_per_s": positions / float(rg["cpu_secs_per_pass_1thread"]),
                    "note": "calculate() includes the reference's synchronous copy of the descent graph, as its own loop does"}
            v, info = reference_cli_best(name)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "reference", "runs": info["all_runs"],
                                    "sample": "UNMODIFIED reference CLI `%s`: wall time minus the wall time of the same command "
                                              "with -b 0 -i 1 (set-up %.2f s)" % (info["command"], info["setup_s"])}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    hst.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: the synthetic pedigree, -R 8 replicates + MC3 coupled chains sharded across 1/2/4/8 GPUs
# ---------------------------------------------------------------------------------------------

C4_REPLICATES, C4_CHAINS, C4_EXCHANGE, C4_STEP_ITERATIONS, C4_SCORING_PERIOD = 8, 2, 10, 20, 10


def run_c4(args, rank, world, local_rank):
    """A FIXED job -- 8 replicates, each a Metropolis-coupled ladder of 2 chains (16 chains) -- over `world` GPUs: strong
    scaling.  Replicate r runs on rank r mod world (swiftlink::ReplicateJob: all of a rank's ladders resident at once,
    swaps are pointer exchanges on the ladder's GPU); every timed step ends with the merge of the per-rank LOD
    accumulators and swap counters over NCCL (swiftlink_b200/dist.py), inside the timed region."""
    import torch
    from swiftlink_b200 import build, capi, dist as sdist, host as H
    build.build()
    if not torch.cuda.is_available() or capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        dist = init_nccl(local_rank)
    paths = workload_files(N_MARKERS, "c4r%d" % rank)
    hst = H.Host(*paths, lodscores=N_LOD)
    assert hst.set_peel_by_names(load_order()["order"]), "committed elimination order rejected"
    ids = sdist.chain_placement(C4_REPLICATES, world)[rank]
    total_steps = args.warmup + args.steps
    t0 = time.perf_counter()
    with quiet_stdout():
        job = H.Job(hst, ids, 0, total_steps * C4_STEP_ITERATIONS, scoring_period=C4_SCORING_PERIOD, seed=20261017, device=local_rank,
                    lsampler_prob=LSAMPLER_PROB, si_iterations=1, mc3_chains=C4_CHAINS, exchange_period=C4_EXCHANGE)
    t_setup = time.perf_counter() - t0
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    merge_ms = []

    def step():
        n = job.advance(C4_STEP_ITERATIONS) if ids else C4_STEP_ITERATIONS
        assert n == C4_STEP_ITERATIONS
        r = job.results()                                  # waits for the rank's chains, reads the raw accumulators
        t1 = time.perf_counter()
        raw, count = sdist.merge_lod(torch.from_numpy(r["raw"]).to(dev), r["count"])
        ok, bad = sdist.merge_swap_stats(r["swap_success"], r["swap_failure"], device=dev)
        torch.cuda.synchronize()
        merge_ms.append(1e3 * (time.perf_counter() - t1))
        return raw, count, ok, bad, r["trait_prob"]
    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    clocks.start()
    merge_ms[:] = []
    secs = 0.0
    out = None
    for _ in range(args.steps):
        barrier()
        ta = time.perf_counter()
        out = step()
        barrier()
        secs += time.perf_counter() - ta
    clock_info = clocks.stop()
    if dist is not None:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    raw, count, ok, bad, tp = out
    value = C4_REPLICATES * C4_STEP_ITERATIONS * args.steps / secs
    lod = sdist.normalise(raw, count, tp)
    n_me = 281
    launches_per_it = 0.5 * 2 + 0.5 * (1 + 2 * ((n_me + 1) // 2))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 200-member consanguineous pedigree, 10k SNPs: a fixed job of %d replicates, each an MC3 ladder "
                                   "of %d chains (exchange every %d iterations), default sampler mix, LOD scoring every %dth iteration, "
                                   "dealt out over the GPUs" % (C4_REPLICATES, C4_CHAINS, C4_EXCHANGE, C4_SCORING_PERIOD),
                       "step": "%d iterations of every replicate (cold-chain iterations are what `value` counts) + the NCCL merge of the "
                               "LOD tables and swap counters" % C4_STEP_ITERATIONS,
                       "replicates": C4_REPLICATES, "chains_per_ladder": C4_CHAINS, "n_members": hst.N, "n_markers": hst.M,
                       "parallelism": "replicate ladders round-robin over %d GPU(s), %d ladder(s) = %d chains resident per GPU"
                                      % (world, len(ids), len(ids) * C4_CHAINS),
                       "l2": "working set of %d chains per GPU exceeds L2" % (len(ids) * C4_CHAINS)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(raw.numel() * 8),
                    "note": "the job runs through the product API (swiftlink::ReplicateJob via the C ABI); per step the raw LOD "
                            "accumulators come back to the host and go through the all-reduces"},
            "gpu_launches": int(args.steps * len(ids) * C4_CHAINS * C4_STEP_ITERATIONS * launches_per_it),
            "clocks": clock_info,
            "derived": {"merge_ms_per_step": float(np.mean(merge_ms)), "setup_s": t_setup, "scoring_passes_merged": count,
                        "swap_success": [int(x) for x in ok.cpu().numpy()], "swap_failure": [int(x) for x in bad.cpu().numpy()],
                        "lod_max": float(lod.max().item()), "chains_total": C4_REPLICATES * C4_CHAINS}}
    job.close()
    hst.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="synth200", choices=["synth200", "east", "loop", "xlinked", "c4"],
                    help="synth200 (default): BASELINE.json configs[2], the configuration the metric is quoted on; "
                         "east / loop / xlinked: the reference's example pedigrees (configs[0], [1], [4]); c4: configs[3], a fixed "
                         "job of 8 replicates x MC3 ladders over --gpus GPUs (strong scaling, NCCL merges inside the timed region)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-step-seconds", type=float, default=10.0,
                    help="--impl reference: size of the bounded sample one step times (seconds of CPU work, approximately)")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="also time this many replicate chains in flight on one GPU (derived.replicates_in_flight; 1 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "c4":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference's MC3 driver is unreachable (linkage_program.cc:169-170) "
                                                                       "and its -R loop is timed by the synth200 configuration"}))
        else:
            run_c4(args, rank, world, local_rank)
    elif args.config != "synth200":
        if args.impl == "reference":
            run_small_reference(args, rank)
        else:
            run_small_ours(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
