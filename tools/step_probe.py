"""Wall-clock of the individual library calls of one resident bench step (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safepy_b200 import _lib, synthetic as syn
from safepy_b200.ordering import kd_order
from safepy_b200.permutations import make_perm_rows

prof = "--profile" in sys.argv
cfg = syn.make_config("C3", 1.0, shuffle=True)
n, m, P = cfg["n"], cfg["m"], 1000
net, attrs = cfg["net"], cfg["attributes"]
nr = cfg["radius"] * (net["x"].max() - net["x"].min())
ctx = _lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
rows = torch.from_numpy(make_perm_rows(attrs, P, 7)).cuda()
b = torch.from_numpy(attrs).cuda()
counts = torch.zeros((2, n, m), dtype=torch.int32, device="cuda")
order = kd_order(net["x"], net["y"])
ctx.profile(prof)
for it in range(6):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    counts.zero_(); t.append(time.perf_counter())
    plan = _lib.Enrichment(nb, b_dev=b.data_ptr(), dtype=np.float32, shape=(n, m)); t.append(time.perf_counter())
    plan.set_node_order(order); t.append(time.perf_counter())
    plan.perm_counts_dev(rows.data_ptr(), P, counts[0].data_ptr(), counts[1].data_ptr(), "sum", "auto"); t.append(time.perf_counter())
    st = plan.stats(); plan.close(); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print("step %d: zero %.1f create %.1f order %.1f perm_counts %.1f close %.1f sync %.1f total %.1f ms" % (
        (it,) + tuple(1e3 * (t[i + 1] - t[i]) for i in range(6)) + (1e3 * (t[-1] - t[0]),)), flush=True)
    if prof:
        print("   kernel ms:", {k: round(ctx.kernel_ms(k)[0], 1) for k in ("gemm", "gather", "fixup", "prep", "score")})
