"""Device time of the z-score permutation null (neighborhood_score_type='z-score') on a named configuration:
tensor-core path (six digit planes per accumulation, comparison in the epilogue) vs the 'sum' null on the same inputs.
    python tools/zscore_probe.py [--config C3] [--perms 48]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from safepy_b200 import _lib, get_context, synthetic as syn  # noqa: E402
from safepy_b200.ordering import kd_order  # noqa: E402
from safepy_b200.permutations import make_perm_rows  # noqa: E402


def main():
    name = sys.argv[sys.argv.index("--config") + 1] if "--config" in sys.argv else "C3"
    perms = int(sys.argv[sys.argv.index("--perms") + 1]) if "--perms" in sys.argv else 48
    ctx = get_context()
    cfg = syn.make_config(name, shuffle=True)
    net, n, m, attrs = cfg["net"], cfg["n"], cfg["m"], cfg["attributes"]
    nr = cfg["radius"] * (np.max(net["x"]) - np.min(net["x"]))
    nb = _lib.Neighborhoods(ctx, n)
    if cfg["metric"] == "euclidean":
        nb.euclid(net["x"], net["y"], nr)
    else:
        nb.shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
    rows = make_perm_rows(attrs, perms, 7)
    dev = torch.device("cuda", ctx.device)
    rows_dev = torch.from_numpy(rows).to(dev)
    attrs_dev = torch.from_numpy(attrs).to(dev)
    counts = torch.zeros((2, n, m), dtype=torch.int32, device=dev)
    out = {"config": name, "n": n, "m": m, "perms": perms}
    for score in ("sum", "z-score"):
        plan = _lib.Enrichment(nb, b_dev=attrs_dev.data_ptr(), dtype=np.float32, shape=(n, m))
        plan.set_node_order(kd_order(net["x"], net["y"]))
        plan.perm_counts_dev(rows_dev.data_ptr(), perms, counts[0].data_ptr(), counts[1].data_ptr(), score, "tc")
        ctx.synchronize()
        counts.zero_()
        torch.cuda.synchronize()
        ctx.profile(True)
        for k in _lib.KERNEL_CLASSES:
            ctx.kernel_ms(k)
        t0 = time.perf_counter()
        plan.perm_counts_dev(rows_dev.data_ptr(), perms, counts[0].data_ptr(), counts[1].data_ptr(), score, "tc")
        ctx.synchronize()
        dt = time.perf_counter() - t0
        kern = {k: round(ctx.kernel_ms(k)[0], 2) for k in ("gemm", "gather", "fixup", "score", "prep")}
        ctx.profile(False)
        st = plan.stats()
        out[score] = {"seconds": dt, "ms_per_permutation": 1e3 * dt / perms, "scores_per_s": float(n) * m * perms / dt,
                      "fixup_fraction": st["fixups"] / max(1, st["fixups"] + st["decided"]), "kernel_ms": kern}
        plan.close()
    out["zscore_over_sum"] = out["z-score"]["seconds"] / out["sum"]["seconds"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
