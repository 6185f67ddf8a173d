import sys, torch
sys.path.insert(0, ".")
from position_induced_transformer_b200.posatt import position_attention
dev = torch.device("cuda:0")
B, N, D, H, batched = 10, 972, 256, 2, True
g = torch.Generator().manual_seed(0)
mesh = torch.rand((B, N, 2) if batched else (N, 2), generator=g).to(dev)
vals = torch.randn(B, N, D, generator=g).to(dev).requires_grad_(True)
scale = (torch.rand(H, generator=g) * 20 + 5).to(dev).requires_grad_(True)
up = torch.randn(B, N, (1 + H) * D, generator=g).to(dev)
for _ in range(3):
    out = position_attention(mesh, mesh, vals, scale, 1.0, "euclid", True)
    out.backward(up)
torch.cuda.synchronize()
