"""CPU study for the next round: what would 16-bit operands / 16-bit accumulators cost the lower-bound pass (pass C)?

The pass computes one Gram value per (catalog row, query prototype) on the tensor cores and rejects a row for a query
when  LB = hull(G~) - margin  is above the optimistic threshold tau_opt, where hull() is the squared distance from the
row to the affine hull of the query's prototypes (a quadratic in the Gram values), G~ the Gram values as the MMA
delivers them and margin = 2 * max_k |G~_k - G_k| bounded rigorously by u * |p|max * |e| (s lies in the simplex).
Every survivor is rescored exactly, so the only price of a cheaper MMA is MORE SURVIVORS.  This script emulates the
roundings of five MMA variants on a bench-like workload (features relu(N(0,1))*10/31.9098 projected by Xavier
weight-norm heads: F=1024, K=3, d=64) and counts survivors at the quantile the kernel works at (512 expected rows of
1 M under the exact distance).  Pure torch float64 on the CPU; prints one JSON line per variant.

    python tools/lb_precision_study.py [N] [Q]
"""
import json
import sys

import torch

torch.manual_seed(633)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 48
F, K, D = 1024, 3, 64
f64 = torch.float64


def xavier(a, b):
    return (torch.rand(a, b, dtype=f64) * 2 - 1) * (6.0 / (a + b)) ** 0.5


def features(n):
    return torch.randn(n, F, dtype=f64).clamp_(min=0) * (10.0 / 31.9098)


def wn(x, V):
    return (x @ V) / V.norm(dim=0)


V0, Vp = xavier(F, D), xavier(F, K * D)
E = torch.cat([wn(features(min(20000, N - lo)), V0) for lo in range(0, N, 20000)])
P = wn(features(Q), Vp).reshape(Q, K, D)
mu = E.mean(0)
E, P = E - mu, P - mu                                     # the kernel works on centred operands


def exact_dist(Pq, E):
    dk = ((E[:, None, :] - Pq[None, :, :]) ** 2).sum(-1)  # [N,K]
    s = torch.softmax(-dk, dim=-1)
    m = s @ Pq
    return ((E - m) ** 2).sum(-1)


def hull_lb(G, e2, Pq):
    """Squared distance from e to the affine hull of the prototypes, from Gram values G[n,k] = p_k.e (float64)."""
    # minimise |e - sum s_k p_k|^2 over sum s = 1:  KKT system [PP 1; 1 0][s; l] = [G; 1]
    PP = Pq @ Pq.T
    M = torch.zeros(K + 1, K + 1, dtype=f64)
    M[:K, :K], M[:K, K], M[K, :K] = PP, 1.0, 1.0
    rhs = torch.cat([G, torch.ones(G.shape[0], 1, dtype=f64)], 1)
    sol = torch.linalg.solve(M, rhs.T).T
    s = sol[:, :K]
    return e2 - 2 * (s * G).sum(-1) + ((s @ PP) * s).sum(-1)


def rnd(x, mant_bits):
    """Round to nearest with `mant_bits` explicit mantissa bits (tf32/fp16: 10, bf16: 7)."""
    m, e = torch.frexp(x)
    scale = 2.0 ** (mant_bits + 1)
    return torch.ldexp(torch.round(m * scale) / scale, e)


def gram(Pq, E, op_bits, acc_bits=None, kstep=16):
    a, b = rnd(E, op_bits), rnd(Pq, op_bits)
    if acc_bits is None:
        return a @ b.T
    acc = torch.zeros(E.shape[0], K, dtype=f64)
    for j in range(0, D, kstep):                          # the accumulator is rounded after every MMA instruction
        acc = rnd(acc + a[:, j:j + kstep] @ b[:, j:j + kstep].T, acc_bits)
    return acc


VARIANTS = {
    # name: (operand mantissa bits, accumulator mantissa bits, rigorous u of one Gram value relative to |p||e|)
    "tf32 operands, fp32 accumulate (fallback planes)": (10, None, 1.1 * 2.0 ** -10),
    "fp16 operands, fp32 accumulate (shipped)": (10, None, 1.1 * 2.0 ** -10),   # same significand as tf32; range is ample
    "bf16 operands, fp32 accumulate": (7, None, 1.1 * 2.0 ** -7),
    "fp16 operands, fp16 accumulate (d/16 roundings)": (10, 10, 1.1 * 2.0 ** -10 + (D // 16) * 2.0 ** -11),
    "bf16 operands, fp16 accumulate": (7, 10, 1.1 * 2.0 ** -7 + (D // 16) * 2.0 ** -11),
}
REL = 2.0e-4                                              # CFL_PLANE_REL: fp32 evaluation margin of the bound
frac = 512.0 / 1.0e6
res = {k: dict(surv=0.0, viol=0, worst=0.0) for k in VARIANTS}
base = 0.0
for q in range(Q):
    Pq = P[q]
    dist = exact_dist(Pq, E)
    tau = torch.quantile(dist, frac)
    base += float((dist <= tau).sum())
    e2 = (E * E).sum(-1)
    pmax = Pq.norm(dim=-1).max()
    G_true = E @ Pq.T
    for name, (ob, ab, u) in VARIANTS.items():
        G = gram(Pq, E, ob, ab)
        margin = 2.0 * u * pmax * e2.sqrt() + REL * e2
        lb = hull_lb(G, e2, Pq) - margin
        r = res[name]
        r["surv"] += float((lb <= tau).sum())
        r["viol"] += int((lb > dist * (1 + 1e-12) + 1e-12).sum())       # a bound above the distance would be a bug
        r["worst"] = max(r["worst"], float(((G - G_true).abs().max(-1).values / (pmax * e2.sqrt())).max()))
for name, r in res.items():
    print(json.dumps(dict(variant=name, survivors_per_query=round(r["surv"] / Q, 1),
                          rows_under_tau_exact=round(base / Q, 1), inflation=round(r["surv"] / base, 2),
                          bound_violations=r["viol"], worst_gram_error_over_pe=float("%.3g" % r["worst"]),
                          rigorous_u=float("%.3g" % VARIANTS[name][2]), N=N, Q=Q)), flush=True)
