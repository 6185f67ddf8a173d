import sys, torch
sys.path.insert(0, '.')
from position_induced_transformer_b200 import workloads, posatt
dev = torch.device('cuda:0')
w = workloads.make_darcy(43, 2).to(dev)
gen = torch.Generator().manual_seed(0)
ins, tgt = w.make_batch(gen, 2)
ins = tuple(x.to(dev) for x in ins); tgt = tgt.to(dev)
orig_key = posatt.rowstat_cache.key
def key(*a):
    k = orig_key(*a); print('key', k[0], k[1], k[2:]); return k
posatt.rowstat_cache.key = key
for i in range(3):
    loss = w.loss(tgt, workloads.run_model(w, ins)); loss.backward()
    print(i, posatt.rowstat_cache.hits, posatt.rowstat_cache.misses, len(posatt.rowstat_cache.entries))
