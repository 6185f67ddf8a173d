"""Candidate-count histogram of the decoder tile plan of a workload (default darcy421): how well sorting by candidate set packs the tiles."""
import collections
import sys

import torch

sys.path.insert(0, ".")
from position_induced_transformer_b200 import posatt as pa, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "darcy421"
dev = torch.device("cuda:0")
w = workloads.WORKLOADS[name](8).to(dev)
mesh = w.meshes[0].reshape(-1, w.model.space_dim)
ltt = w.model.mesh_ltt
values = torch.zeros(8, ltt.shape[0], 64, device=dev)
variant = w.model.up._variant
mo, mi, st, period, stats, _ = pa.prepare_meshes(mesh, ltt, values, w.model.n_head, variant, w.model.de_local)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
plan = pa.build_tail_plan(st, mo, mi, period, stats)
torch.cuda.synchronize()
print(f"{name}: N={st.N} M={st.M} tiles={plan.n_tiles} plan build {1e3 * (time.perf_counter() - t0):.2f} ms")
sizes = plan.tile_cnt.cpu().tolist()
hist = collections.Counter(sizes)
tot = len(sizes)
print("candidates per 32-row tile: mean %.2f" % (sum(sizes) / tot))
for k in sorted(hist):
    print(f"  {k:4d}: {hist[k]:6d}  {100.0 * hist[k] / tot:5.1f} %")
print("<= 8: %.1f %%   <= 16: %.1f %%" % (100.0 * sum(v for k, v in hist.items() if k <= 8) / tot, 100.0 * sum(v for k, v in hist.items() if k <= 16) / tot))
# per-row candidate counts for reference
# load balance of the backward's contiguous partition (one CTA per SM, ceil(tiles / SMs) tiles each): k-steps of 8 candidates per CTA
import math
sms = torch.cuda.get_device_properties(dev).multi_processor_count
per = math.ceil(tot / sms)
ks = [(c + 7) // 8 for c in sizes]
loads = [sum(ks[i:i + per]) for i in range(0, tot, per)]
print(f"backward partition: {len(loads)} CTAs x {per} tiles; k-steps per CTA min {min(loads)} mean {sum(loads) / len(loads):.1f} max {max(loads)}"
      f" (max / mean = {max(loads) * len(loads) / sum(loads):.2f}); tiles in the last CTA: {tot - per * (len(loads) - 1)}")
