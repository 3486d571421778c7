"""Generates the committed golden fixtures (tests/golden/*.npz) from the fp64 oracle.

    python tests/golden/make_golden.py        # rewrites the .npz files

The reference has no golden vectors and cannot be imported here (no TensorFlow),
so these are oracle-generated ("parity unpinned" -- see oracle/cfl_oracle.py).  Inputs are
stored in float32 (what the CUDA path consumes); expected outputs are the fp64 oracle
evaluated on those float32 inputs.  Seed 633 = the reference's default (cfl/utils.py:89).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cfl_oracle as O  # noqa: E402

f32 = np.float32
f64 = np.float64


def _pair_inputs(rng, B, K, d, kind):
    v = rng.normal(size=(B, d))
    if kind == "plain":
        P = v[:, None, :] + rng.normal(size=(B, K, d))
    elif kind == "offset":                      # large common offset (App. B row 2)
        v = v + 10
        P = v[:, None, :] + rng.normal(size=(B, K, d))
    elif kind == "neardup":                     # dist << |v|^2
        P = rng.normal(size=(B, K, d))
        P[:, 0, :] = v + 1e-2 * rng.normal(size=(B, d))
    elif kind == "far":                         # peaked softmax / large distances
        P = 6 * rng.normal(size=(B, K, d))
    return v.astype(f32), P.astype(f32)


def case_pair_pcd():
    rng = np.random.default_rng(633)
    out = {}
    cfgs = [(1, 8, "plain"), (2, 10, "plain"), (3, 64, "plain"), (4, 20, "plain"), (5, 12, "far"),
            (4, 15, "offset"), (3, 64, "neardup"), (8, 128, "plain"), (4, 10, "neardup")]
    for i, (K, d, kind) in enumerate(cfgs):
        B = 17 if i % 2 else 24
        v, P = _pair_inputs(rng, B, K, d, kind)
        up = rng.normal(size=B).astype(f32)
        dist, s, dk = O.pcd_dist(v.astype(f64), P.astype(f64), return_aux=True)
        dv, dP = O.pcd_dist_bwd(v.astype(f64), P.astype(f64), up.astype(f64))
        out.update({f"c{i}_v": v, f"c{i}_P": P, f"c{i}_up": up, f"c{i}_dist": dist,
                    f"c{i}_s": s, f"c{i}_dv": dv, f"c{i}_dP": dP})
    return out


def case_pair_modes():
    rng = np.random.default_rng(634)
    B, K, d = 48, 4, 20
    a = rng.normal(size=(B, d)).astype(f32)
    b = rng.normal(size=(B, d)).astype(f32)
    Pt = rng.normal(size=(B, K, d)).astype(f32)
    w = O.softmax(rng.normal(size=(B, K))).astype(f32)
    up = rng.normal(size=B).astype(f32)
    da, dPt, dw = O.monomer_dist_bwd(a.astype(f64), Pt.astype(f64), w.astype(f64), up.astype(f64))
    return dict(a=a, b=b, Pt=Pt, w=w, up=up,
                monomer=O.monomer_dist(a.astype(f64), Pt.astype(f64), w.astype(f64)),
                siamese=O.siamese_dist(a.astype(f64), b.astype(f64)),
                da=da, dPt=dPt, dw=dw)


LOSS_OPTS = [
    dict(theta=1e-6),
    dict(theta=0.8, pos_weight=0.0625),
    dict(theta=1e-7, pos_weight=0.25, lambda_m=0.5),
    dict(theta=1.5, use_threshold=False, caffe_margin=2.0),
    dict(theta=-1.0, pos_weight=0.5, caffe_margin=1.5),
]


def case_loss():
    rng = np.random.default_rng(635)
    Bp, Bn = 100, 100
    dp = (np.abs(rng.normal(size=Bp)) * 1.5).astype(f32)
    dn = (np.abs(rng.normal(size=Bn)) * 3.0).astype(f32)
    dn[:3] = dp[:3]                             # some ties across the two sets
    out = dict(d_pos=dp, d_neg=dn)
    for i, o in enumerate(LOSS_OPTS):
        o = dict(o)
        th = f64(f32(o.pop("theta")))
        L = O.dist_losses(dp.astype(f64), dn.astype(f64), th, reg=0.0, **o)
        gp, gn, gth = O.dist_losses_bwd(dp.astype(f64), dn.astype(f64), th, **o)
        for k in ("p_loss_pos", "p_loss_neg", "thres_loss", "cd_loss", "total_loss", "accuracy",
                  "margins", "pos_dists_adapt", "neg_dists_adapt", "score_pos", "score_neg"):
            out[f"o{i}_{k}"] = np.asarray(L[k])
        out[f"o{i}_gpos"], out[f"o{i}_gneg"], out[f"o{i}_gtheta"] = gp, gn, np.asarray(gth)
    return out


def case_project():
    rng = np.random.default_rng(636)
    B, F, K, d = 40, 96, 3, 12
    x = np.maximum(rng.normal(size=(B, F)), 0).astype(f32) * 10
    V0 = O.xavier_uniform(rng, F, d)
    Vp = O.xavier_uniform(rng, F, K * d)
    g0 = rng.uniform(0.5, 1.5, d).astype(f32)
    gp = rng.uniform(0.5, 1.5, K * d).astype(f32)
    b0 = (0.1 * rng.normal(size=d)).astype(f32)
    bp = (0.1 * rng.normal(size=K * d)).astype(f32)
    in_scale = f32(1 / 31.9098)
    xs = x.astype(f64) * f64(in_scale)
    out = dict(x=x, V0=V0, Vp=Vp, g0=g0, gp=gp, b0=b0, bp=bp, in_scale=np.asarray(in_scale))
    for act in ("linear", "tanh", "sigmoid", "relu"):
        r = O.build_prototypes(xs, {"outputs": (V0.astype(f64), g0.astype(f64), b0.astype(f64)),
                                    "prototype_outputs": (Vp.astype(f64), gp.astype(f64), bp.astype(f64))},
                               "pcd", K, d, act=act)
        out[f"e_{act}"] = r["activations"]
        out[f"P_{act}"] = r["prototype_activations"]
    lat, pcd = O.fcencoder(xs, V0.astype(f64), b0.astype(f64), Vp.astype(f64), bp.astype(f64), K, d)
    out["plain_e"], out["plain_P"] = lat, pcd
    dy = rng.normal(size=(B, K * d)).astype(f32)
    _, _, z = O.fc_weight_norm(xs, Vp.astype(f64), gp.astype(f64), bp.astype(f64), None, return_pre=True)
    dV, dg, db = O.fc_weight_norm_bwd(xs, Vp.astype(f64), gp.astype(f64), z, dy.astype(f64))
    out.update(dy=dy, dVp=dV, dgp=dg, dbp=db)
    return out


def case_rank():
    rng = np.random.default_rng(637)
    Q, K, d, N, k = 12, 3, 16, 700, 20
    E = rng.normal(size=(N, d)).astype(f32)
    E[100] = E[7]                               # exact duplicate rows -> exact ties
    E[650] = E[7]
    anchors = E[rng.integers(0, N, Q)]
    Pq = (anchors[:, None, :] + 0.5 * rng.normal(size=(Q, K, d))).astype(f32)
    D = O.all_pairs_dist(Pq.astype(f64), E.astype(f64))
    vals, idx = O.rank_topk(D, k)
    return dict(E=E, Pq=Pq, dist=D, top_val=vals, top_idx=idx, k=np.asarray(k))


def case_rank_monomer():
    """Monomer mode on the cross product (base.py:109-117 on every (source, target) pair) with exact duplicate rows."""
    rng = np.random.default_rng(639)
    Q, K, d, N, k = 10, 3, 12, 600, 20
    Pt = rng.normal(size=(N, K, d)).astype(f32)
    Pt[7] = Pt[7, 1][None, :]                   # all prototypes of row 7 coincide ...
    Pt[90] = Pt[7]                              # ... and rows 90, 555 duplicate it: exact ties
    Pt[555] = Pt[7]
    a = (Pt[rng.integers(0, N, Q), rng.integers(0, K, Q)] + 0.5 * rng.normal(size=(Q, d))).astype(f32)
    a[:4] = (Pt[7, 1] + 0.01 * rng.normal(size=(4, d))).astype(f32)
    w = O.softmax(2.0 * rng.normal(size=(Q, K))).astype(f32)
    D = O.all_pairs_monomer_dist(a.astype(f64), w.astype(f64), Pt.astype(f64))
    vals, idx = O.rank_topk(D, k)
    return dict(a=a, w=w, Pt=Pt, dist=D, top_val=vals, top_idx=idx, k=np.asarray(k))


def case_auc():
    rng = np.random.default_rng(638)
    n = 5000
    scores = np.round(rng.normal(size=n), 2).astype(f32)    # many ties
    labels = (rng.uniform(size=n) < 0.06).astype(np.uint8)
    scores[labels == 1] += f32(0.5)
    two_u, npos, nneg = O.auc_exact(scores, labels)
    cp = int((scores[labels == 1] > 0).sum())
    cn = int((scores[labels == 0] <= 0).sum())
    return dict(scores=scores, labels=labels, two_u=np.asarray(two_u, dtype=np.int64),
                n_pos=np.asarray(npos), n_neg=np.asarray(nneg), correct=np.asarray(cp + cn))


CASES = dict(pair_pcd=case_pair_pcd, pair_modes=case_pair_modes, loss=case_loss,
             project=case_project, rank=case_rank, rank_monomer=case_rank_monomer, auc=case_auc)

if __name__ == "__main__":
    for name, fn in CASES.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **fn())
        print(name, os.path.getsize(path), "bytes")
