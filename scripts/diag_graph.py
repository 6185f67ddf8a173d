import sys, copy, torch
sys.path.insert(0, '.')
from position_induced_transformer_b200 import workloads
from position_induced_transformer_b200.data_parallel import FlatGradients
from position_induced_transformer_b200.graphed import GraphedTrainStep
dev = torch.device('cuda:0')
gen = torch.Generator().manual_seed(5)
w = workloads.make_darcy(43, batch=2).to(dev)
batches = [w.make_batch(gen, 2) for _ in range(3)]
batches = [(tuple(x.to(dev) for x in ins), tgt.to(dev)) for ins, tgt in batches]
ref_model = copy.deepcopy(w.model); ref_model.mesh_ltt = w.model.mesh_ltt
def loss_of(model):
    return lambda ins, tgt: w.loss(tgt, model(w.meshes[0], ins[0], w.meshes[0]))
opt_g = torch.optim.SGD(w.model.parameters(), lr=1e-5)
step = GraphedTrainStep(list(w.model.parameters()), loss_of(w.model), opt_g, batches[0][0], batches[0][1], warmup=2)
opt_e = torch.optim.SGD(ref_model.parameters(), lr=1e-5)
flat = FlatGradients(ref_model.parameters(), 1)
def eager_step(ins, tgt):
    flat.zero(); loss = loss_of(ref_model)(ins, tgt); loss.backward(); opt_e.step(); return loss
for _ in range(2): print('warm', float(eager_step(*batches[0])))
d = max(float((a-b).abs().max()) for a, b in zip(w.model.parameters(), ref_model.parameters()))
print('param diff after warmup', d)
for ins, tgt in batches:
    le = float(eager_step(ins, tgt)); lg = float(step(ins, tgt))
    d = max(float((a-b).abs().max()) for a, b in zip(w.model.parameters(), ref_model.parameters()))
    print('eager', le, 'graph', lg, 'param diff', d)
