import sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden, t, rel_linf
from position_induced_transformer_b200 import workloads
import position_induced_transformer_b200.pit as P
dev = torch.device('cuda:0')
torch.set_float32_matmul_precision('highest')
g = load_golden('model_darcy43')
ctor = {k[5:]: g[k] for k in g if k.startswith('ctor/')}
args = [int(ctor[k]) for k in ("space_dim", "in_dim", "out_dim", "hid_dim", "n_head", "n_blocks")]
model = workloads.DarcyPiT(*args, t(ctor['mesh_ltt'], dev), 0.02, 0.02).to(dev)
model.load_state_dict({k[6:]: t(v) for k, v in g.items() if k.startswith('param/')})
for k, v in model.named_parameters():
    if k.endswith('lmda'):
        a = P.head_scale(v.detach()).cpu().flatten(); b = P.head_scale(v.detach().cpu()).flatten()
        print(k, a.tolist(), b.tolist(), (a.view(torch.int32) - b.view(torch.int32)).tolist())
ins = [t(g[f'input/{i}'], dev) for i in range(3)]
out = model(*ins)
print('device scale err', rel_linf(out.detach().cpu(), t(g['out'])))
orig = P.head_scale
P.head_scale = lambda l: orig(l.cpu()).to(l.device)
out = model(*ins)
print('cpu scale err', rel_linf(out.detach().cpu(), t(g['out'])))
