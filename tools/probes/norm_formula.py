"""Which fp32 formula does torch.norm(c[:, None] - c, dim=2, p=2) evaluate on this GPU?  (probe for the kNN-graph kernel)"""
import torch
dev = "cuda"
for n in (20, 100, 200, 500, 1000):
    g = torch.Generator().manual_seed(n)
    c = torch.rand((n, 2), generator=g).to(dev)
    ref = torch.norm(c[:, None] - c, dim=2, p=2)
    d = c[:, None] - c
    dx, dy = d[..., 0].contiguous(), d[..., 1].contiguous()
    a, b = dx * dx, dy * dy
    cand1 = torch.sqrt(a + b)                                             # fl(fl(dx^2) + fl(dy^2))
    cand2 = torch.sqrt((dy.double() * dy.double() + a.double()).float())  # fma(dy, dy, fl(dx^2))
    cand3 = torch.sqrt((dx.double() * dx.double() + b.double()).float())  # fma(dx, dx, fl(dy^2))
    cand4 = torch.sqrt((dx.double() ** 2 + dy.double() ** 2)).float()     # exact sum, one rounding, then sqrt in f64
    cd = torch.cdist(c, c)
    print(n, "mismatches: mul+mul+add", int((ref != cand1).sum()), " fma(dy,dy,dx2)", int((ref != cand2).sum()),
          " fma(dx,dx,dy2)", int((ref != cand3).sum()), " f64", int((ref != cand4).sum()), " cdist", int((ref != cd).sum()))
    # batched form [B, n, 2]
    cb = torch.rand((4, n, 2), generator=g).to(dev)
    refb = torch.norm(cb[:, :, None] - cb[:, None], dim=3, p=2)
    db = cb[:, :, None] - cb[:, None]
    print("   batched mul+mul+add", int((refb != torch.sqrt(db[..., 0] * db[..., 0] + db[..., 1] * db[..., 1])).sum()))
    # ties inside top-k: how torch orders equal values
    vals, idx = torch.topk(ref, k=min(10, n - 1), dim=1, largest=False)
    ties = (vals[:, 1:] == vals[:, :-1])
    print("   rows with a tie inside the top-k:", int(ties.any(dim=1).sum()), " of them index-ascending:",
          int((ties & (idx[:, 1:] > idx[:, :-1])).sum()), "descending:", int((ties & (idx[:, 1:] < idx[:, :-1])).sum()))
# forced ties: integer grid coordinates
c = torch.stack(torch.meshgrid(torch.arange(8.), torch.arange(8.), indexing="ij"), dim=-1).reshape(-1, 2).to(dev) / 8
ref = torch.norm(c[:, None] - c, dim=2, p=2); ref[torch.arange(64), torch.arange(64)] = 1e9
vals, idx = torch.topk(ref, k=12, dim=1, largest=False)
ties = (vals[:, 1:] == vals[:, :-1])
print("grid: tie pairs", int(ties.sum()), "index-ascending", int((ties & (idx[:, 1:] > idx[:, :-1])).sum()), "descending",
      int((ties & (idx[:, 1:] < idx[:, :-1])).sum()))
print(idx[0].tolist(), [round(v, 4) for v in vals[0].tolist()])
print(idx[27].tolist(), [round(v, 4) for v in vals[27].tolist()])
