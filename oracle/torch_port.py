"""Op-for-op torch-CPU port of the reference's TF graph for the hot path.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/cfl_oracle.py header for the import rule).

Two uses:
* fp64 + autograd: an independent transcription that pins the closed forms in
  ``cfl_oracle`` (tests/test_oracle.py);
* fp32 with all host threads: the ``cpu_baseline`` / ``--impl reference`` arm of
  ``bench.py``.  TensorFlow cannot be installed here, so this is the reference CPU path
  ``kind: "port"``: the same op sequence the TF graph executes, including the
  materialised ``[B,K,d]`` broadcast tensors (cfl/models/base.py:129-137).
"""
from __future__ import annotations

import torch


def fc_weight_norm(x, V, g=None, b=None):
    """cfl/layers.py:80-90."""
    out = x @ V
    scaler = (g if g is not None else 1.0) / torch.sqrt(torch.sum(V * V, dim=0))
    out = scaler.reshape(1, -1) * out
    if b is not None:
        out = out + b
    return out


def fc_plain(x, W, b):
    """cfl/models/dist.py:45-65."""
    return x @ W + b


def pcd_dist(v, P):
    """cfl/models/base.py:125-146, materialising the [B,K,d] tensors like the TF graph."""
    B, K, d = P.shape
    if K > 1:
        diff = v.reshape(-1, 1, d) - P
        logits = -torch.sum(diff * diff, dim=-1)
        scales = torch.softmax(logits, dim=-1)
        means = torch.sum(P * scales.reshape(-1, K, 1), dim=-2)
        return torch.sum((v - means) ** 2, dim=-1)
    diff = v - P.reshape(-1, d)
    return torch.sum(diff * diff, dim=-1)


def monomer_dist(a, Pt, w):
    """cfl/models/base.py:109-117."""
    d = a.shape[-1]
    diff = a.reshape(-1, 1, d) - Pt
    return torch.sum(w * torch.sum(diff * diff, dim=-1), dim=-1)


def siamese_dist(a, b):
    """cfl/models/base.py:119-123."""
    return torch.sum((a - b) ** 2, dim=-1)


def thresholder(dist, theta):
    """cfl/models/blocks.py:21-22; clamp(min=) routes the tie gradient like tf.maximum."""
    return -1.0 * dist + torch.clamp(theta, min=1e-6)


def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def dist_total_loss(d_pos, d_neg, theta, pos_weight=None, use_threshold=True,
                    caffe_margin=None, lambda_m=None, reg=0.0):
    """cfl/models/cfl.py:868-929."""
    sp = thresholder(d_pos, theta)
    sn = thresholder(d_neg, theta)
    lp = sigmoid_ce(sp, torch.ones_like(sp)).mean()
    ln = sigmoid_ce(sn, torch.zeros_like(sn)).mean()
    thres = lp * pos_weight + ln if pos_weight else lp + ln
    total = torch.as_tensor(reg, dtype=d_pos.dtype)
    if use_threshold:
        total = total + thres
    if caffe_margin:
        cp = d_pos.mean()
        if pos_weight:
            cp = cp * pos_weight
        cn = torch.clamp(caffe_margin - d_neg, min=0).mean()
        total = total + 0.5 * (cp + cn)
    elif lambda_m:
        cp = d_pos.mean() * lambda_m
        if pos_weight:
            cp = cp * pos_weight
        total = total + cp
    return total, lp, ln


def score_pairs_reference(xs, xt, V0, g0, b0, Vp, gp, bp, K, d, theta, in_scale=1.0):
    """One eval batch the way cfl/utils.py:245 drives it: normalise, encode source
    (prototype head) and target (e0 head), build_dist, threshold."""
    P = fc_weight_norm(xs * in_scale, Vp, gp, bp).reshape(-1, K, d)
    v = fc_weight_norm(xt * in_scale, V0, g0, b0)
    return thresholder(pcd_dist(v, P), theta)


def all_pairs_scores_reference(Pq, E, theta, pair_batch=500):
    """The only way the reference could rank a catalog: the QxN cross product pushed
    through its pair scorer in batches (cfl/bin/predict.py:195 batch size 500).  Here
    each query's prototypes are broadcast against ``pair_batch`` candidates at a time."""
    Q, K, d = Pq.shape
    N = E.shape[0]
    out = torch.empty(Q, N, dtype=E.dtype)
    for q in range(Q):
        for c0 in range(0, N, pair_batch):
            v = E[c0:c0 + pair_batch]
            P = Pq[q:q + 1].expand(v.shape[0], K, d)
            out[q, c0:c0 + pair_batch] = thresholder(pcd_dist(v, P), theta)
    return out


def all_pairs_scores_blocked(Pq, E, theta, cand_block=65536):
    """Stronger CPU line: same op sequence as ``pcd_dist`` but with large [q, block]
    batches so MKL/ATen threads are busy (still direct-difference form, [b,K,d] tensors)."""
    Q, K, d = Pq.shape
    N = E.shape[0]
    out = torch.empty(Q, N, dtype=E.dtype)
    for q in range(Q):
        for c0 in range(0, N, cand_block):
            v = E[c0:c0 + cand_block]
            P = Pq[q:q + 1].expand(v.shape[0], K, d)
            out[q, c0:c0 + cand_block] = thresholder(pcd_dist(v, P), theta)
    return out


def all_pairs_scores_gram(Pq, E, theta):
    """Strongest CPU line (not what the reference does): Gram form through MKL SGEMM."""
    Q, K, d = Pq.shape
    G = (Pq.reshape(Q * K, d) @ E.T).reshape(Q, K, -1)           # [Q,K,N]
    e2 = (E * E).sum(-1)                                          # [N]
    p2 = (Pq * Pq).sum(-1)                                        # [Q,K]
    if K == 1:
        dist = e2[None, :] + p2[:, 0:1] - 2 * G[:, 0, :]
    else:
        dk = e2[None, None, :] + p2[:, :, None] - 2 * G
        s = torch.softmax(-dk, dim=1)
        pp = torch.einsum("qkd,qld->qkl", Pq, Pq)
        dist = e2[None, :] - 2 * (s * G).sum(1) + torch.einsum("qkn,qkl,qln->qn", s, pp, s)
    return thresholder(dist, theta)
