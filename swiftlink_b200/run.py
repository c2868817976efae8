"""`swift` for one 8 x B200 box: the `-R` replicates (and their MC3 ladders) of one linkage job dealt out over the GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m swiftlink_b200.run -p ped -m map -d dat [-R 8] [-M -z 3 -y 10] [-b B -i I -x X -l P -n 5 -s S] [-X] [-o out]

One process per GPU (a single process without torchrun).  Replicate r -- LinkageProgram::run_pedigree's loop,
linkage_program.cc:96-108 -- runs on rank r mod world, all of a rank's replicates resident on its GPU at once
(swiftlink::ReplicateJob, csrc/host/job.cc); an MC3 ladder (mc3.cc:81-200) stays on its replicate's GPU, where a swap
is a pointer exchange.  Nothing travels while the chains sample; at the end the per-rank RAW log-sum LOD accumulators,
the scoring-pass counts and the ladders' swap counters are combined with small NCCL all-reduces over NVLink
(swiftlink_b200/dist.py: LODscores::merge_results, lod_score.h:98-105, is an element-wise log-sum plus a sum of counts,
i.e. MAX / SUM-of-exp / SUM all-reduces) and rank 0 writes the reference's output file (linkage_writer.cc:14-92).
A chain's draws are keyed by (seed, chain id, iteration): the merged table does not depend on the number of GPUs
beyond floating-point summation order.
"""
import argparse
import os
import sys
import time

import numpy as np


def parse(argv=None):
    ap = argparse.ArgumentParser(prog="python -m swiftlink_b200.run", add_help=True)
    ap.add_argument("-p", "--pedigree", required=True)
    ap.add_argument("-m", "--map", required=True)
    ap.add_argument("-d", "--dat", required=True)
    ap.add_argument("-o", "--output", default="swiftlink.out")
    ap.add_argument("-i", "--iterations", type=int, default=50000)
    ap.add_argument("-b", "--burnin", type=int, default=50000)
    ap.add_argument("-s", "--sequentialimputation", type=int, default=1000)
    ap.add_argument("-x", "--scoringperiod", type=int, default=10)
    ap.add_argument("-l", "--lsamplerprobability", type=float, default=0.5)
    ap.add_argument("-n", "--lodscores", type=int, default=5)
    ap.add_argument("-R", "--runs", type=int, default=1)
    ap.add_argument("-M", "--mcmcmc", action="store_true")
    ap.add_argument("-z", "--chains", type=int, default=1)
    ap.add_argument("-y", "--exchangeperiod", type=int, default=10)
    ap.add_argument("-t", "--temperatures", default=None, help="comma-separated, one per chain")
    ap.add_argument("-X", "--sexlinked", action="store_true")
    ap.add_argument("-q", "--peelseqiter", type=int, default=1000000)
    ap.add_argument("-S", "--seed", type=int, default=20261017)
    ap.add_argument("--peel-order-json", default=None, help="elimination order as person ids (skips the peel search)")
    ap.add_argument("--backend", default=None, help="torch.distributed backend (default: nccl)")
    return ap.parse_args(argv)


def run_job(hst, args, rank, world, device, dist_group=None, step=None, on_step=None):
    """the rank's share of the job: returns (merged normalised LOD table [M-1][n], info) on every rank"""
    import torch
    from swiftlink_b200 import dist as sdist, host as H
    ids = sdist.chain_placement(args.runs, world)[rank]
    temps = None if not args.temperatures else [float(x) for x in args.temperatures.split(",")]
    mc3 = args.chains if args.mcmcmc else 1
    t0 = time.perf_counter()
    job = H.Job(hst, ids, args.burnin, args.iterations, scoring_period=args.scoringperiod, seed=args.seed, device=device,
                lsampler_prob=args.lsamplerprobability, si_iterations=args.sequentialimputation, mc3_chains=mc3,
                exchange_period=args.exchangeperiod, temperatures=temps)
    t_setup = time.perf_counter() - t0
    total = args.burnin + args.iterations
    done = 0
    while done < total:
        n = job.advance(total - done if step is None else min(step, total - done))
        if n <= 0:
            break
        done += n
        if on_step:
            on_step(done)
    r = job.results()
    t_run = time.perf_counter() - t0 - t_setup
    # ---- merge over the ranks (NCCL all-reduces; no-ops in a single process) ----
    dev = torch.device("cuda", device) if torch.cuda.is_available() and (args.backend or "nccl") == "nccl" else torch.device("cpu")
    t1 = time.perf_counter()
    raw, count = sdist.merge_lod(torch.from_numpy(r["raw"]).to(dev), r["count"], group=dist_group)
    ok, bad = sdist.merge_swap_stats(r["swap_success"], r["swap_failure"], group=dist_group, device=dev)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t_merge = time.perf_counter() - t1
    tp = r["trait_prob"]
    if world > 1:
        # every rank computed the same ln P(T); ranks without a replicate have none
        import torch.distributed as dist
        t = torch.tensor([tp if ids else -1e300], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=dist_group)
        tp = float(t.item())
    lod = sdist.normalise(raw, count, tp).cpu().numpy().reshape(hst.M - 1, hst.nlod)
    job.close()
    return lod, dict(replicates=ids, setup_s=t_setup, run_s=t_run, merge_s=t_merge, scoring_passes=count,
                     swap_success=ok.cpu().numpy(), swap_failure=bad.cpu().numpy(), trait_prob=tp)


def main(argv=None):
    args = parse(argv)
    import torch
    from swiftlink_b200 import build, capi, host as H
    build.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() == 0:
        raise SystemExit("swiftlink_b200.run: no CUDA device -- the product path has no CPU fallback")
    local_rank = local_rank % capi.device_count()          # (several ranks may share a device: tests on a one-GPU box, gloo)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        backend = args.backend or "nccl"
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank) if backend == "nccl" else None)
    hst = H.Host(args.pedigree, args.map, args.dat, sex_linked=args.sexlinked, lodscores=args.lodscores)
    if args.peel_order_json:
        import json
        with open(args.peel_order_json) as f:
            assert hst.set_peel_by_names(json.load(f)["order"]), "elimination order rejected"
    else:
        hst.build_peel(args.peelseqiter, seed=args.seed)          # same seed on every rank: the same sequence
    lod, info = run_job(hst, args, rank, world, local_rank)
    if rank == 0:
        hst.write_results(args.output, lod)
        its = args.runs * (args.burnin + args.iterations)
        print("%d replicate(s) x %d iterations on %d GPU(s): %.2f s sampling (%.1f iterations/s), merge %.1f ms, %d scoring passes"
              % (args.runs, args.burnin + args.iterations, world, info["run_s"], its / max(info["run_s"], 1e-9),
                 1e3 * info["merge_s"], info["scoring_passes"]))
        if args.mcmcmc and args.chains > 1:
            for i in range(args.chains - 1):
                s, f = int(info["swap_success"][i]), int(info["swap_failure"][i])
                print("%d -- %d : %.3f (%d/%d)" % (i, i + 1, s / max(s + f, 1), s, s + f))
        print("wrote %s" % args.output)
    hst.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
