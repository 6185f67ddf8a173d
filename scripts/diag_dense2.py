import sys, torch, ctypes as C
sys.path.insert(0, '.')
from position_induced_transformer_b200 import _cabi, posatt
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(1)
N, M, D, B, H = 128, 32, 64, 1, 1
mo = torch.rand(N, 2, generator=g).to(dev); mi = torch.rand(M, 2, generator=g).to(dev)
vals = torch.ones(B, M, D, device=dev)
scale = torch.tensor([1.0], device=dev)
st = posatt._Stage(mo, mi, vals, H, 'euclid')
v_min, v_lo, v_hi, w, masked = posatt.row_statistics(st, mo, mi, None, 1.0)
out = torch.full((B, N, H * D), -7.0, device=dev)
rowsum = torch.full((H, N), -7.0, device=dev)
ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
rs = posatt._rowstat_struct(v_min, v_lo, v_hi, w, masked)
rc = _cabi.lib.pit_posatt_forward(C.byref(st.problem), mo.data_ptr(), mi.data_ptr(), None, vals.data_ptr(), scale.data_ptr(), C.byref(rs),
                                  out.data_ptr(), H * D, 0, 0, rowsum.data_ptr(), ws.data_ptr(), 1 << 20, torch.cuda.current_stream().cuda_stream)
print('rc', rc, _cabi.lib.pit_last_error())
torch.cuda.synchronize()
print('rowsum', rowsum[0, :6].tolist())
print('out row0', out[0, 0, :8].tolist(), 'row 127', out[0, 127, :8].tolist())
print('out stats', float(out.min()), float(out.max()))
d2 = ((mo[:, None, :] - mi[None, :, :]) ** 2).sum(-1)
print('expected rowsum', torch.exp(d2.min(1, keepdim=True).values - d2).sum(1)[:6].tolist())
