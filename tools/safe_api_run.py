"""Wall clock of the SAFE class itself (BASELINE.json's second metric: define_neighborhoods + compute_pvalues seconds)
on a named synthetic configuration, on one GPU or under torchrun (one rank per GPU).

    python tools/safe_api_run.py --config C3
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/safe_api_run.py --config C5

Times host call to host return of the two methods (graph and attribute objects already built, second call), prints one
JSON line on rank 0.  With several ranks only rank 0 receives the [N, M] result arrays (sf.results_rank = 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--perms", type=int, default=1000)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--outputs", default=None,
                    help="comma-separated sf.host_outputs (default: all five result arrays, as upstream); several sets "
                         "separated by ';' are timed one after the other")
    ap.add_argument("--how", default="randomization", help="compute_pvalues(how=...): randomization | auto | hypergeometric")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from safepy_b200 import SAFE, synthetic as syn
    cfg = syn.make_config(args.config, args.scale, shuffle=True)
    net, attrs = cfg["net"], cfg["attributes"]
    sf = SAFE(verbose=False, device=local_rank)
    t0 = time.perf_counter()
    sf.load_network(edges=net["edges"] if len(net["edges"]) else None, x=net["x"], y=net["y"],
                    length=net["length"] if len(net["edges"]) else None)
    t_load = time.perf_counter() - t0
    sf.node_distance_metric = cfg["metric"]
    sf.neighborhood_radius = cfg["radius"]
    sf.random_seed = 7
    sf.results_rank = 0
    sf.assume_graph_unchanged = True     # the graph object is not touched between load_network and the timed calls
    sf.load_attributes(attribute_file=attrs)
    # one JSON line per set of host outputs (sets separated by ';'), all in this process
    for outputs in (args.outputs.split(";") if args.outputs else [None]):
        if outputs:
            sf.host_outputs = tuple(outputs.split(","))
        runs = []
        for _ in range(args.repeats):
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
            sf.define_neighborhoods()
            t1 = time.perf_counter()
            sf.compute_pvalues(how=args.how, num_permutations=args.perms)
            t2 = time.perf_counter()
            runs.append(dict(define_neighborhoods_s=t1 - t0, compute_pvalues_s=t2 - t1,
                             phases=getattr(sf, "last_enrichment_seconds", None)))
        if dist:
            import torch
            worst = torch.tensor([runs[-1]["define_neighborhoods_s"], runs[-1]["compute_pvalues_s"]], device="cuda")
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            worst = [float(v) for v in worst.cpu()]
        else:
            worst = [runs[-1]["define_neighborhoods_s"], runs[-1]["compute_pvalues_s"]]
        if rank == 0:
            n, m = attrs.shape
            print(json.dumps({
                "metric": "define_neighborhoods+compute_pvalues sec", "config": args.config, "n": n, "m": m,
                "perms": args.perms, "how": args.how, "n_gpus": world, "host_outputs": list(sf.host_outputs), "node_distance_metric": cfg["metric"],
                "define_neighborhoods_s": worst[0], "compute_pvalues_s": worst[1], "total_s": worst[0] + worst[1],
                "timing": "host wall clock of the last of %d calls, max over ranks" % args.repeats,
                "rank0_runs": runs, "load_network_s": t_load,
                "mean_neighborhood": float(np.mean(np.sum(sf.neighborhoods, axis=1))),
                "enriched_cells": float(np.nansum(sf.nes_binary)) if sf.nes_binary is not None else None,
                "gemm_stats_rank0": getattr(sf, "last_enrichment_stats", None)}))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
