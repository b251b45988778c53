import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pcops_b200 import sa_modules as sam
torch.manual_seed(0)
for rows, K, N in ((32, 128, 256), (32, 64, 64), (256, 64, 128)):
    x = torch.zeros(rows, K, device="cuda"); dy = torch.zeros(rows, N, device="cuda")
    x[0, 0] = 1.0; x[1, 5] = 2.0; x[9, 33] = 3.0
    dy[0, 0] = 1.0; dy[1, 7] = 1.0; dy[9, 40] = 1.0; dy[0, 1] = 5.0
    dw, db = sam.dense_weight_grad(x, dy)
    want = x.double().t() @ dy.double()
    nz = dw.nonzero().tolist()
    print(rows, K, N, "nonzeros got:", [(i, j, round(dw[i, j].item(), 3)) for i, j in nz][:12], " want:", [(i, j, want[i, j].item()) for i, j in want.nonzero().tolist()])
    print("   db nonzero:", [(i, round(db[i].item(), 3)) for i in db.nonzero().flatten().tolist()][:8])
x = torch.randn(64, 128, device="cuda"); dy = torch.randn(64, 256, device="cuda")
dw, db = sam.dense_weight_grad(x, dy)
want = (x.double().t() @ dy.double())
print("random: max|got|", dw.abs().max().item(), "max|want|", want.abs().max().item(), "err", (dw.double() - want).abs().max().item())
