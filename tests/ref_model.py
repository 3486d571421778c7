"""fp64 reference of one CFL / Dist train step, built from the op-for-op torch port
(oracle/torch_port.py) + autograd + the oracle's TF-style Adam.  Test infrastructure."""
import numpy as np
import torch

from oracle import cfl_oracle as O
from oracle import torch_port as T

ACT = {None: lambda t: t, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu}


def head(x, p, weight_norm, act):
    y = T.fc_weight_norm(x, p["V"], p.get("g"), p.get("b")) if weight_norm else x @ p["V"] + p["b"]
    return y


def pair_dists(cfg, Wsrc, Wdst, xs, xt):
    K, d, act = cfg["K"], cfg["d"], ACT[cfg.get("act")]
    wn = cfg["weight_norm"]
    xs, xt = xs * cfg["in_scale"], xt * cfg["in_scale"]
    if cfg["dist_type"] == "pcd":
        P = act(head(xs, Wsrc["proto"], wn, None)).reshape(-1, K, d)
        v = act(head(xt, Wdst["e0"], wn, None))
        return T.pcd_dist(v, P)
    if cfg["dist_type"] == "siamese":
        return T.siamese_dist(act(head(xs, Wsrc["e0"], wn, None)), act(head(xt, Wdst["e0"], wn, None)))
    pre = head(xs, Wsrc["e0"], wn, None)
    a = act(pre)
    w = torch.softmax(T.fc_weight_norm(pre, Wsrc["gate"]["V"], Wsrc["gate"]["g"], None), dim=-1)
    Pt = act(head(xt, Wdst["proto"], wn, None)).reshape(-1, K, d)
    return T.monomer_dist(a, Pt, w)


def train_step(cfg, weights, theta, batch, adam_state, step):
    """weights: {'src': {'e0': {V,g,b}, 'proto': {...}, 'gate': {...}}, 'dst': same-or-other}; numpy fp64.
    Returns (losses dict, new weights, new theta, new adam_state)."""
    tw = {}
    leaves = []
    for enc, hs in weights.items():
        tw[enc] = {}
        for hn, ps in hs.items():
            tw[enc][hn] = {}
            for pn, arr in ps.items():
                t = torch.tensor(arr, dtype=torch.float64, requires_grad=True)
                tw[enc][hn][pn] = t
                leaves.append(((enc, hn, pn), t))
    if cfg.get("shared", True):
        tw["dst"] = tw["src"]
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    xsp, xtp, xsn, xtn = (torch.tensor(b, dtype=torch.float64) for b in batch)
    dp = pair_dists(cfg, tw["src"], tw["dst"], xsp, xtp)
    dn = pair_dists(cfg, tw["src"], tw["dst"], xsn, xtn)
    reg = 0.0
    if cfg.get("reg_const"):
        seen = set()
        for (enc, hn, pn), t in leaves:
            if pn in ("V", "b") and id(t) not in seen:
                reg = reg + 0.5 * cfg["reg_const"] * (t ** 2).sum()
                seen.add(id(t))
    total, lp, ln = T.dist_total_loss(dp, dn, th, pos_weight=cfg.get("pos_weight"),
                                      use_threshold=cfg.get("use_threshold", True),
                                      caffe_margin=cfg.get("caffe_margin"), lambda_m=cfg.get("lambda_m"), reg=0.0)
    total = total + reg
    params = [t for _, t in leaves]
    grads = torch.autograd.grad(total, params + [th], allow_unused=True)
    pw = cfg.get("pos_weight")
    thres = lp * pw + ln if pw else lp + ln
    if not cfg.get("use_threshold", True):
        gth = torch.autograd.grad(thres, th, allow_unused=True)[0]
    else:
        gth = grads[-1]
    new_w = {enc: {hn: {} for hn in hs} for enc, hs in weights.items()}
    new_state = {}
    for ((enc, hn, pn), t), g in zip(leaves, grads[:-1]):
        g = np.zeros_like(t.detach().numpy()) if g is None else g.numpy()
        m, v = adam_state.get((enc, hn, pn), (np.zeros_like(g), np.zeros_like(g)))
        p, m, v = O.adam_tf(t.detach().numpy(), g, m, v, step, cfg["lr"], cfg.get("beta1", 0.9), cfg.get("beta2", 0.999))
        new_w[enc][hn][pn] = p
        new_state[(enc, hn, pn)] = (m, v)
    gth = 0.0 if gth is None else float(gth)
    m, v = adam_state.get("theta", (np.zeros(()), np.zeros(())))
    nth, m, v = O.adam_tf(np.asarray(theta, dtype=np.float64), np.asarray(gth), m, v, step, cfg["lr"],
                          cfg.get("beta1", 0.9), cfg.get("beta2", 0.999))
    new_state["theta"] = (m, v)
    sp, sn = (T.thresholder(dp, th), T.thresholder(dn, th))
    acc = 0.5 * (float((sp > 0).double().mean()) + float((sn <= 0).double().mean()))
    return dict(total=float(total), lp=float(lp), ln=float(ln), thres=float(thres), acc=acc, dp=dp.detach().numpy(),
                dn=dn.detach().numpy()), new_w, float(nth), new_state
