import sys, os, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/compatibility-family-learning_b200')
import bench
from cfl import _native as nat
dev=torch.device('cuda',0)
d,K5,Q5,N5=20,4,1024,2_000_000
g = torch.Generator(device=dev).manual_seed(bench.SEED + 50 + d)
E = torch.randn(N5, d, generator=g, device=dev)
gq = torch.Generator(device=dev).manual_seed(bench.SEED + 51 + d)
anchors = torch.randn(Q5, d, generator=gq, device=dev)
Pq = anchors[:, None, :] + 0.5 * torch.randn(Q5, K5, d, generator=gq, device=dev)
mu = E.mean(0)
img = nat.catalog_pack(E, K5, mu)
tv, ti, st, thr = nat.score_topk(Pq, E, 100, mu=mu, image=img, want_stats="thresholds")
print(dict(zip(nat.SCORE_STAT_NAMES, st.tolist())))
tau, tau_opt, redo = thr[0], thr[1], thr[2]
bad = torch.nonzero(redo > -3e38).flatten()
print('redo queries', bad.tolist(), 'tau', tau[bad].tolist(), 'tau_opt', tau_opt[bad].tolist())
for q in bad.tolist():
    P = Pq[q] - mu
    dd = torch.cdist(P, P)**2
    print('pairwise proto dist^2', dd)
    D = ((E[:, None, :] - Pq[q][None])**2).sum(-1)
    s = torch.softmax(-D, 1)
    m = s @ Pq[q]
    dist = ((E - m)**2).sum(-1)
    print('rows under tau_opt', int((dist <= tau_opt[q]).sum()), 'under tau', int((dist <= tau[q]).sum()))
