import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from hitgeom._lib import lib, ptr, stream_ptr, check
B, N = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
x = torch.randn(B, N, 3, device="cuda")
x = x / x.norm(dim=-1).amax(dim=1)[:, None, None]
y = x + 0.01 * torch.randn_like(x)
L = lib()
nb = L.hg_nn_bidir_workspace_bytes(B, N, N, 3)
ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
m1 = torch.empty(B, N, device="cuda"); m2 = torch.empty(B, N, device="cuda")
a1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); a2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
check(L.hg_nn_bidir_f32(ptr(x), ptr(y), B, N, N, 3, ptr(m1), ptr(a1), ptr(m2), ptr(a2), ptr(ws), nb, stream_ptr()), "nn")
torch.cuda.synchronize()
al = lambda v: (v + 255) // 256 * 256
o = 0
colres = ws[o:o + B * N * 8].view(torch.int64).cpu().numpy().view(np.uint64); o += al(B * N * 8)
rowres = ws[o:o + B * N * 8].view(torch.int64).cpu().numpy().view(np.uint64); o += al(B * N * 8)
colsec = ws[o:o + B * N * 4].view(torch.int32).cpu().numpy().view(np.uint32); o += al(B * N * 4)
rowsec = ws[o:o + B * N * 4].view(torch.int32).cpu().numpy().view(np.uint32); o += al(B * N * 4)
eps2 = ws[o:o + B * 4].view(torch.float32).cpu().numpy(); o += al(B * 4)
amb = ws[o:o + (B * 2 * N + 1) * 4].view(torch.int32).cpu().numpy()
def unord(u):
    u = u.astype(np.uint32)
    r = np.where(u & 0x80000000, u ^ 0x80000000, ~u).astype(np.uint32)
    return r.view(np.float32)
cnt = amb[0]
ent = amb[1:1 + cnt]
e = ent % (2 * N)
print("eps2", eps2[:3], "ambiguous", cnt, "of", B * 2 * N, "cols", (e < N).sum(), "rows", (e >= N).sum())
cb, cs = unord((colres >> np.uint64(32)).astype(np.uint32)), unord(colsec)
rb, rs = unord((rowres >> np.uint64(32)).astype(np.uint32)), unord(rowsec)
flag = ((rowres & np.uint64(0x80000000)) != 0)
print("col: best", cb[:5], "sec", cs[:5], "gap<eps", (cs - cb <= eps2[0]).mean())
print("row: best", rb[:5], "sec", rs[:5], "gap<eps", (rs - rb <= eps2[0]).mean(), "flag", flag.mean())
print("exact m1", m1.flatten()[:5].cpu().numpy(), "m2", m2.flatten()[:5].cpu().numpy())
