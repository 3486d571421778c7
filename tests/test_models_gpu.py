"""Model-level parity on the GPU: the reference-shaped constructors and the fused train step
(projection -> paired distance + loss -> backward -> TF-style Adam) against an fp64 torch-autograd
restatement of the reference graph (tests/ref_model.py), same seeded weights and batches."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_model as R  # noqa: E402
from oracle import cfl_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _weights_of(model):
    def h(head):
        d = {"V": _np(head.V)}
        if head.g is not None:
            d["g"] = _np(head.g)
        if head.b is not None:
            d["b"] = _np(head.b)
        return d
    out = {}
    for name, enc in (("src", model.enc_src), ("dst", model.enc_dst)):
        out[name] = {hn: h(getattr(enc, hn)) for hn in ("e0", "proto", "gate") if getattr(enc, hn) is not None}
    return out


def _batch(rng, B, F):
    return [np.maximum(rng.normal(size=(B, F)), 0).astype(np.float32) * 3 for _ in range(4)]


CASES = [
    dict(model="dist", F=64, K=4, d=10, norm=58.388599, reg_const=0.0),
    dict(model="dist", F=96, K=3, d=20, norm=2.0, reg_const=1e-3),
    dict(model="cfl", F=128, K=3, d=64, dist_type="pcd", pos_weight=0.0625, data_norm=31.9098),
    dict(model="cfl", F=80, K=4, d=20, dist_type="pcd", pos_weight=0.25, act_type="tanh", reg_const=5e-4, lambda_m=0.5),
    dict(model="cfl", F=80, K=2, d=12, dist_type="pcd", directed=True, act_type="sigmoid"),
    dict(model="cfl", F=72, K=3, d=16, dist_type="monomer", act_type="relu", reg_const=1e-3),
    dict(model="cfl", F=72, K=1, d=16, dist_type="siamese", use_threshold=False, caffe_margin=2.0, pos_weight=0.5),
    dict(model="cfl", F=72, K=1, d=24, dist_type="siamese", use_threshold=True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items() if k in ("model", "dist_type", "K", "d", "act_type")))
def test_train_steps_match_reference_graph(case):
    from cfl import ops, variables as vs
    from cfl.models.cfl import CFL
    from cfl.models.dist import Dist
    vs.reset_default_graph()
    vs.set_seed(633)
    F, K, d = case["F"], case["K"], case["d"]
    lr = 1e-2
    if case["model"] == "dist":
        model = Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=50, lr=lr, beta1=0.9,
                     beta2=0.999, normalize_value=case["norm"], data_normalizer=ops.normalizer(case["norm"], 0.0),
                     reg_const=case["reg_const"])
        cfg = dict(K=K, d=d, dist_type="pcd", weight_norm=False, in_scale=1.0 / case["norm"], lr=lr,
                   reg_const=case["reg_const"], shared=True)
    else:
        dn = case.get("data_norm")
        norm = ops.normalizer_v2((F,), norm=dn) if dn else None
        model = CFL(input_shape=(F,), batch_size=50, latent_size=d, num_components=K, model_type="linear",
                    dist_type=case["dist_type"], act_type=case.get("act_type"), data_type="linear",
                    use_threshold=case.get("use_threshold", True), pos_weight=case.get("pos_weight"),
                    caffe_margin=case.get("caffe_margin"), lambda_m=case.get("lambda_m"),
                    reg_const=case.get("reg_const", 0.0), directed=case.get("directed", False), lr=lr,
                    data_normalizer=norm, data_norm=dn)
        cfg = dict(K=K if case["dist_type"] != "siamese" else 1, d=d, dist_type=case["dist_type"], weight_norm=True,
                   in_scale=(1.0 / dn) if dn else 1.0, lr=lr, act=case.get("act_type"),
                   pos_weight=case.get("pos_weight"), use_threshold=case.get("use_threshold", True),
                   caffe_margin=case.get("caffe_margin"), lambda_m=case.get("lambda_m"),
                   reg_const=case.get("reg_const", 0.0), shared=not case.get("directed", False))
    # move theta off the clamp so both branches of the loss are exercised, and perturb g / biases
    rng = np.random.default_rng(7)
    with torch.no_grad():
        model.raw_threshold.fill_(0.8)
        for p in model._params:
            if p.dim() == 1:
                p.add_(torch.as_tensor(0.1 * rng.normal(size=p.shape), dtype=torch.float32, device=p.device))
    weights = _weights_of(model)
    if cfg["shared"]:
        weights = {"src": weights["src"]}
    theta, state = 0.8, {}
    B = 50
    for step in range(1, 4):
        batch = _batch(rng, B, F)
        out = model.train_step(*[torch.as_tensor(b).cuda() for b in batch])
        ref, weights, theta, state = R.train_step(cfg, weights, theta, batch, state, step)
        np.testing.assert_allclose(out["s_p_loss_pos"], ref["lp"], rtol=1e-4)
        np.testing.assert_allclose(out["s_p_loss_neg"], ref["ln"], rtol=1e-4)
        np.testing.assert_allclose(out["s_total_loss"], ref["total"], rtol=1e-4)
        np.testing.assert_allclose(out["s_accuracy"], ref["acc"], rtol=1e-12)
        np.testing.assert_allclose(_np(model.s_pos_dists)[:, 0], ref["dp"], rtol=1e-4)
        got = _weights_of(model)
        for enc, hs in weights.items():
            for hn, ps in hs.items():
                for pn, arr in ps.items():
                    # Adam's first steps move every weight by ~lr regardless of the gradient's size,
                    # so compare the UPDATE: tolerance relative to lr
                    np.testing.assert_allclose(got[enc][hn][pn], arr, atol=2e-3 * lr * step + 1e-7, rtol=0,
                                               err_msg=f"step {step} {enc}/{hn}/{pn}")
        np.testing.assert_allclose(float(model.raw_threshold), theta, atol=2e-3 * lr * step + 1e-7)


def test_c2_shape_train_step_parity():
    """BASELINE config 2 at its real shape (monomer run.sh: F=4096 fc7-like features, K=4, d=20, B=100,
    normalize 58.388599): distances and losses of the fused step within 1e-4 of the fp64 graph even
    though the projection runs as 3xTF32 on tensor cores."""
    from cfl import ops, variables as vs
    from cfl.models.dist import Dist
    vs.reset_default_graph()
    vs.set_seed(633)
    F, K, d, B, lr = 4096, 4, 20, 100, 1e-3
    model = Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=B, lr=lr, beta1=0.9, beta2=0.999,
                 normalize_value=58.388599, data_normalizer=ops.normalizer(58.388599, 0.0))
    cfg = dict(K=K, d=d, dist_type="pcd", weight_norm=False, in_scale=1.0 / 58.388599, lr=lr, reg_const=0.0, shared=True)
    rng = np.random.default_rng(2)
    with torch.no_grad():
        model.raw_threshold.fill_(0.3)
    weights = {"src": _weights_of(model)["src"]}
    theta, state = 0.3, {}
    for step in range(1, 3):
        batch = [np.minimum(np.maximum(rng.normal(size=(B, F)), 0) * 20, 58.388599).astype(np.float32) for _ in range(4)]
        batch[1] = (batch[0] * 0.97 + 0.03 * batch[1]).astype(np.float32)      # positives: close pairs
        out = model.train_step(*[torch.as_tensor(b).cuda() for b in batch])
        ref, weights, theta, state = R.train_step(cfg, weights, theta, batch, state, step)
        np.testing.assert_allclose(_np(model.s_pos_dists)[:, 0], ref["dp"], rtol=1e-4)
        np.testing.assert_allclose(_np(model.s_neg_dists)[:, 0], ref["dn"], rtol=1e-4)
        np.testing.assert_allclose(out["s_total_loss"], ref["total"], rtol=1e-4)
        np.testing.assert_allclose(out["s_accuracy"], ref["acc"], rtol=1e-12)


def test_variable_names_follow_the_reference_scopes():
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.models.dist import Dist
    vs.reset_default_graph()
    CFL(input_shape=(32,), latent_size=8, num_components=2, model_type="linear", dist_type="monomer",
        use_threshold=True, directed=True)
    names = set(vs.all_variables())
    for n in ("CFL/DistEncoderSrc/outputs/fully_connected/V", "CFL/DistEncoderSrc/outputs/fully_connected/g",
              "CFL/DistEncoderDst/prototype_outputs/fully_connected/V",
              "CFL/DistEncoderSrc/monomer_outputs/fully_connected/g", "CFL/Thresholder/threshold/threshold"):
        assert n in names, n
    assert not any(n.endswith("biases") for n in names), "monomer encoders have no biases (base.py:45-46)"
    vs.reset_default_graph()
    Dist(input_shape=(32,), latent_size=8, num_components=2, batch_size=10, lr=1e-3, beta1=0.9, beta2=0.999)
    names = set(vs.all_variables())
    for n in ("Dist/Encoder/latent_outputs/fully_connected/weights", "Dist/Encoder/pcd_outputs/fully_connected/biases",
              "Dist/Thresholder/threshold/threshold"):
        assert n in names, n


def test_eager_constructors_and_autograd_match_oracle():
    """FCPCD / Thresholder used the reference way (build_dist on two encoder applications), with
    gradients flowing through the kernels via autograd."""
    from cfl import variables as vs
    from cfl.models.blocks import FCPCD, Thresholder
    vs.reset_default_graph()
    vs.set_seed(1)
    rng = np.random.default_rng(3)
    B, F, K, d = 40, 48, 3, 12
    xs = torch.as_tensor(rng.normal(size=(B, F)).astype(np.float32)).cuda()
    xt = torch.as_tensor(rng.normal(size=(B, F)).astype(np.float32)).cuda()
    src = FCPCD(xs, num_outputs=d, num_components=K, input_shape=(F,), batch_size=B, activation_fn="tanh", name="enc")
    tgt = FCPCD(xt, num_outputs=d, num_components=K, input_shape=(F,), batch_size=B, activation_fn="tanh", name="enc", reuse=True)
    dist = src.build_dist(tgt)
    pred = Thresholder(dist)
    assert dist.shape == (B, 1) and pred.outputs.shape == (B, 1)
    V0 = vs.all_variables()["enc/outputs/fully_connected/V"]
    Vp = vs.all_variables()["enc/prototype_outputs/fully_connected/V"]
    loss = torch.nn.functional.softplus(-pred.outputs).mean()
    gV0, gVp, gth = torch.autograd.grad(loss, [V0, Vp, pred.raw_threshold])
    # oracle
    p = {k: _np(v) for k, v in vs.all_variables().items()}
    params = {"outputs": (p["enc/outputs/fully_connected/V"], p["enc/outputs/fully_connected/g"], p["enc/outputs/fully_connected/biases"]),
              "prototype_outputs": (p["enc/prototype_outputs/fully_connected/V"], p["enc/prototype_outputs/fully_connected/g"],
                                    p["enc/prototype_outputs/fully_connected/biases"])}
    so = O.build_prototypes(_np(xs), params, "pcd", K, d, act="tanh")
    to = O.build_prototypes(_np(xt), params, "pcd", K, d, act="tanh")
    do = O.pcd_dist(to["activations"], so["prototype_activations"])
    np.testing.assert_allclose(_np(dist)[:, 0], do, rtol=1e-4)
    tV0 = torch.tensor(params["outputs"][0], requires_grad=True)
    tVp = torch.tensor(params["prototype_outputs"][0], requires_grad=True)
    from oracle import torch_port as T
    P = torch.tanh(T.fc_weight_norm(torch.tensor(_np(xs)), tVp, torch.tensor(params["prototype_outputs"][1]),
                                    torch.tensor(params["prototype_outputs"][2]))).reshape(-1, K, d)
    v = torch.tanh(T.fc_weight_norm(torch.tensor(_np(xt)), tV0, torch.tensor(params["outputs"][1]),
                                    torch.tensor(params["outputs"][2])))
    th = torch.tensor(1e-6, dtype=torch.float64, requires_grad=True)
    lo = torch.nn.functional.softplus(-T.thresholder(T.pcd_dist(v, P), th)).mean()
    rV0, rVp, rth = torch.autograd.grad(lo, [tV0, tVp, th])
    np.testing.assert_allclose(_np(gV0), rV0.numpy(), atol=3e-5 * np.abs(rV0.numpy()).max())
    np.testing.assert_allclose(_np(gVp), rVp.numpy(), atol=3e-5 * np.abs(rVp.numpy()).max())
    np.testing.assert_allclose(float(gth), float(rth), rtol=1e-4)


class _FakeData:
    """The slice of SemiDataSet dist_eval / dist_predict use (cfl/input_data.py:503-540)."""

    def __init__(self, feats, pairs_pos, pairs_neg):
        self.feats, self.pairs_pos, self.pairs_neg = feats, pairs_pos, pairs_neg
        self.index_to_asins = ["A%09d" % i for i in range(len(feats))]

    def whole_pos_batches(self, bs):
        for i in range(0, len(self.pairs_pos), bs):
            p = self.pairs_pos[i:i + bs]
            yield self.feats[p[:, 0]], self.feats[p[:, 1]]

    def whole_neg_batches(self, bs):
        for i in range(0, len(self.pairs_neg), bs):
            p = self.pairs_neg[i:i + bs]
            yield self.feats[p[:, 0]], self.feats[p[:, 1]]


def test_dist_eval_and_predict_match_sklearn(tmp_path):
    from sklearn.metrics import roc_auc_score
    from cfl import ops, utils, variables as vs
    from cfl.models.dist import Dist
    vs.reset_default_graph()
    vs.set_seed(5)
    rng = np.random.default_rng(5)
    F = 64
    feats = np.maximum(rng.normal(size=(300, F)), 0).astype(np.float32) * 20
    model = Dist(input_shape=(F,), latent_size=10, num_components=4, batch_size=100, lr=1e-3, beta1=0.9, beta2=0.999,
                 normalize_value=58.388599, data_normalizer=ops.normalizer(58.388599, 0.0))
    with torch.no_grad():
        model.raw_threshold.fill_(0.05)
    data = _FakeData(feats, rng.integers(0, 300, size=(777, 2)), rng.integers(0, 300, size=(1501, 2)))
    rep = utils.dist_eval(None, model, 500, data)
    sp = np.concatenate([_np(model.predict(*b))[:, 0] for b in data.whole_pos_batches(500)])
    sn = np.concatenate([_np(model.predict(*b))[:, 0] for b in data.whole_neg_batches(500)])
    y = np.r_[np.ones(len(sp)), np.zeros(len(sn))]
    s32 = np.r_[sp, sn].astype(np.float32)
    assert abs(rep.auc - roc_auc_score(y, s32)) < 1e-12
    assert rep.two_u == O.auc_exact(s32, y)[0]
    correct = int((sp > 0).sum()) + int((sn <= 0).sum())
    assert rep.accuracy == correct / len(y) and rep.error == 1 - rep.accuracy or abs(rep.error - (1 - rep.accuracy)) < 1e-15
    utils.dist_predict(None, model, data, 500, str(tmp_path), "predict.txt")
    lines = open(tmp_path / "predict.txt").read().splitlines()
    assert len(lines) == 777 + 1501
    a, m, b, sc = lines[0].split()
    assert m == "match" and a == data.index_to_asins[data.pairs_pos[0, 0]] and abs(float(sc) - sp[0]) < 1e-6


def test_conv_model_trains_on_synthetic_mnist_pairs():
    """BASELINE config 1 plumbing: ConvPCD trunk (torch) + kernel heads, a few Adam steps reduce the loss."""
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    vs.reset_default_graph()
    vs.set_seed(633)
    rng = np.random.default_rng(633)
    model = CFL(input_shape=(28, 28, 1), batch_size=100, latent_size=20, num_components=2, model_type="conv",
                dist_type="pcd", use_threshold=True, reg_const=5e-4, lr=1e-3)
    x = rng.uniform(size=(4, 100, 784)).astype(np.float32)
    x[1] = x[0] * 0.9 + 0.05          # positives: near copies; negatives: unrelated
    losses = [model.train_step(*[torch.as_tensor(b).cuda() for b in x])["s_total_loss"] for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert "CFL/DistEncoder/conv1/Conv/V" in vs.all_variables()
    assert model.predict(torch.as_tensor(x[0]).cuda(), torch.as_tensor(x[1]).cuda()).shape == (100, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("K,act", [(2, None), (3, "tanh")])
def test_conv_model_scores_match_the_oracle(K, act):
    """BASELINE config 1 (fashion_30, ConvPCD, d = 60 / (K + 1)): the conv trunk (cfl/models/blocks.py:563-590,
    cfl/layers.py:147-184: weight-normalised 5x5 stride-2 convolutions, SAME padding, lrelu), the weight-norm heads,
    the pcd distance and the Thresholder against the fp64 oracle on the same variables, 1e-4 relative."""
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from oracle import cfl_oracle as O
    vs.reset_default_graph()
    vs.set_seed(633)
    rng = np.random.default_rng(17 + K)
    d = 60 // (K + 1)
    model = CFL(input_shape=(28, 28, 1), batch_size=64, latent_size=d, num_components=K, model_type="conv",
                dist_type="pcd", use_threshold=True, reg_const=5e-4, lr=1e-3, act_type=act)
    xs = rng.uniform(size=(64, 784)).astype(np.float32)
    xt = (0.7 * xs + 0.3 * rng.uniform(size=(64, 784))).astype(np.float32)
    got = model.predict(torch.as_tensor(xs).cuda(), torch.as_tensor(xt).cuda()).reshape(-1).cpu().numpy()
    # perturb the variables away from their initial values (g = 1, b = 0) and predict again
    named = vs.get_collection(model.name)
    with torch.no_grad():
        for k_, v in named.items():
            if k_.endswith("/g"):
                v.mul_(torch.as_tensor(rng.uniform(0.7, 1.3, tuple(v.shape)), dtype=v.dtype, device=v.device))
            elif k_.endswith("/biases"):
                v.add_(torch.as_tensor(rng.normal(size=tuple(v.shape)) * 0.05, dtype=v.dtype, device=v.device))
            elif k_.endswith("threshold"):
                v.fill_(0.8)
    got2 = model.predict(torch.as_tensor(xs).cuda(), torch.as_tensor(xt).cuda()).reshape(-1).cpu().numpy()
    P = {k_: v.detach().cpu().numpy().astype(np.float64) for k_, v in named.items()}
    enc = "CFL/DistEncoder/"
    convs = [(P[enc + "conv%d/Conv/V" % i], P[enc + "conv%d/Conv/g" % i], P[enc + "conv%d/Conv/biases" % i]) for i in (1, 2)]
    params = {"outputs": tuple(P[enc + "outputs/fully_connected/" + n] for n in ("V", "g", "biases")),
              "prototype_outputs": tuple(P[enc + "prototype_outputs/fully_connected/" + n] for n in ("V", "g", "biases"))}

    def oracle_scores(a, b):
        fa = O.conv_pcd_trunk(a.astype(np.float64), convs)
        fb = O.conv_pcd_trunk(b.astype(np.float64), convs)
        src = O.build_prototypes(fa, params, "pcd", K, d, act)
        dst = O.build_prototypes(fb, params, "pcd", K, d, act)
        dist = O.pcd_dist(dst["activations"], src["prototype_activations"])         # target's e0, source's prototypes
        return O.thresholder(dist, P["CFL/Thresholder/threshold/threshold"]).reshape(-1), dist

    want2, dist2 = oracle_scores(xs, xt)
    np.testing.assert_allclose(got2, want2, rtol=1e-4, atol=1e-4 * float(np.abs(dist2).max()))
    assert np.abs(got - got2).max() > 1e-3            # the perturbation mattered: both states were really evaluated


# ---- fixtures produced by executing the reference's own graph code (tests/golden/ref_*.npz) ---------
import ast as _ast
import glob as _glob

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_REF_CFL = sorted(_glob.glob(os.path.join(_GOLDEN, "ref_cfl_*.npz")))
_REF_DIST = sorted(_glob.glob(os.path.join(_GOLDEN, "ref_dist_*.npz")))


def _check_against_reference_fixture(model, z, prefix, loss_keys):
    from cfl import variables as vs
    named = vs.get_collection(model.name)
    sd = {k: torch.tensor(z["var:" + k], dtype=torch.float32) for k in named}
    model.load_state_dict(sd)
    before = {k: v.detach().clone() for k, v in named.items()}
    c = lambda a: torch.tensor(a, dtype=torch.float32).cuda()
    batch = [c(z["in_" + n]) for n in ("pos_source", "pos_target", "neg_source", "neg_target")]
    val = [c(z["in_val_" + n]) for n in ("pos_source", "pos_target", "neg_source", "neg_target")]
    # Tolerances = the contract's 1e-4 relative, each taken relative to the quantity the fp32
    # error is proportional to: distances and losses to themselves; scores (theta+ - dist) to the distance scale they
    # are a difference of; gradients to the largest entry of their tensor (norm-wise: an entry that cancels to ~0
    # carries the absolute error of the sums it is made of).
    dscale = float(max(np.abs(z["out_s_pos_dists"]).max(), np.abs(z["out_s_neg_dists"]).max(), 1e-6))
    # scores of the evaluation node, before the step moves the weights
    for lab, b in (("pos", batch[:2]), ("neg", batch[2:])):
        np.testing.assert_allclose(model.predict(*b).reshape(-1).cpu().numpy(), z["out_s_%s_predicts" % lab][:, 0],
                                   rtol=1e-4, atol=1e-4 * dscale)
    out = model.train_step(*batch, val_batches=None)
    np.testing.assert_allclose(model.s_pos_dists.reshape(-1).cpu().numpy(), z["out_s_pos_dists"][:, 0], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(model.s_neg_dists.reshape(-1).cpu().numpy(), z["out_s_neg_dists"][:, 0], rtol=1e-4, atol=1e-7)
    for mine, ref in loss_keys:
        np.testing.assert_allclose(out[mine], float(z["out_" + ref]), rtol=1e-4, atol=1e-6, err_msg=ref)
    ref_grads = {k.split(":", 1)[1]: z[k] for k in z.files if k.startswith("grad_")}
    assert set(ref_grads) == set(named)
    for k, g in ref_grads.items():
        mine = model._grads[id(named[k])].cpu().numpy()
        scale = max(np.abs(g).max(), 1e-12)
        np.testing.assert_allclose(mine, g, rtol=0, atol=1e-4 * scale + 1e-9, err_msg=k)
        if not np.any(g):      # heads the loss never reads must not move (TF skips None gradients)
            assert torch.equal(named[k].detach(), before[k]), k
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("path", _REF_CFL, ids=[os.path.basename(p)[8:-4] for p in _REF_CFL])
def test_cfl_model_matches_reference_graph_fixture(path):
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.ops import dist_normalizer
    z = np.load(path)
    cfg = _ast.literal_eval(str(z["cfg"]))
    vs.reset_default_graph()
    norms = dist_normalizer(input_shape=(cfg["F"],), ae_shape=None, data_scale=None, data_mean=None,
                            data_norm=cfg["data_norm"], latent_norm=None, data_type="linear")
    m = CFL(input_shape=(cfg["F"],), batch_size=cfg["B"], latent_size=cfg["d"], num_components=cfg["K"],
            model_type="linear", dist_type=cfg["dist_type"], act_type=cfg["act_type"], data_type="linear",
            use_threshold=cfg["use_threshold"], pos_weight=cfg["pos_weight"], caffe_margin=cfg["caffe_margin"],
            lambda_m=cfg["lambda_m"], reg_const=cfg["reg_const"], directed=cfg["directed"], data_normalizer=norms[0],
            data_norm=cfg["data_norm"])
    keys = [("s_p_loss_pos", "s_p_loss_pos"), ("s_p_loss_neg", "s_p_loss_neg"), ("s_thres_loss", "s_thres_loss"),
            ("s_total_loss", "s_total_loss"), ("s_loss_reg", "s_loss_reg"), ("s_accuracy", "s_accuracy"),
            ("s_margins", "s_margins"), ("s_pos_dists_adapt", "s_pos_dists_adapt"),
            ("s_neg_dists_adapt", "s_neg_dists_adapt"), ("s_margin_adapt", "s_margin_adapt")]
    if cfg["caffe_margin"] or cfg["lambda_m"]:
        keys.append(("s_cd_loss", "s_cd_loss"))
    _check_against_reference_fixture(m, z, "CFL", keys)
    c = lambda a: torch.tensor(a, dtype=torch.float32).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("path", _REF_DIST, ids=[os.path.basename(p)[9:-4] for p in _REF_DIST])
def test_dist_model_matches_reference_graph_fixture(path):
    from cfl import variables as vs
    from cfl.models.dist import Dist
    from cfl.ops import normalizer
    z = np.load(path)
    cfg = _ast.literal_eval(str(z["cfg"]))
    vs.reset_default_graph()
    m = Dist(input_shape=(cfg["F"],), latent_size=cfg["d"], num_components=cfg["K"], batch_size=cfg["B"], lr=1e-3,
             beta1=0.9, beta2=0.999, normalize_value=cfg["normalize_value"],
             data_normalizer=normalizer(cfg["normalize_value"], 0.0), reg_const=cfg["reg_const"])
    _check_against_reference_fixture(m, z, "Dist", [("s_p_loss_pos", "s_p_loss_pos"), ("s_p_loss_neg", "s_p_loss_neg"),
                                                    ("s_thres_loss", "thres_loss"), ("s_total_loss", "s_total_loss"),
                                                    ("s_accuracy", "s_accuracy"), ("s_margins", "s_margins")])


@pytest.mark.gpu
def test_dist_eval_and_predict_file_match_reference_utils(tmp_path):
    """tests/golden/ref_eval_predict.npz: the reference's cfl.utils.dist_eval / dist_predict run over its
    own SemiDataSet and CFL graph (eager stand-in).  Ours: same weights, same directory, CUDA scoring
    node + cfl_auc."""
    from cfl import input_data as I
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.ops import dist_normalizer
    from cfl.utils import dist_eval, dist_predict
    z = np.load(os.path.join(_GOLDEN, "ref_eval_predict.npz"))
    F, d, K, B = (int(z[k]) for k in ("F", "d", "K", "B"))
    ids = [str(s) for s in z["ids"]]
    dd = str(tmp_path)
    I.write_features(os.path.join(dd, "features.b"), ids, z["feats"])
    for name, pairs in (("pairs_pos.txt", z["pos"]), ("pairs_neg.txt", z["neg"])):
        with open(os.path.join(dd, name), "w") as f:
            for a, b in pairs:
                f.write("%s match %s\n" % (ids[a], ids[b]))
    vs.reset_default_graph()
    data_norm = tuple(float(x) for x in z["data_norm"])
    norms = dist_normalizer(input_shape=(F,), ae_shape=None, data_scale=None, data_mean=None, data_norm=data_norm,
                            latent_norm=None, data_type="linear")
    m = CFL(input_shape=(F,), batch_size=B, latent_size=d, num_components=K, model_type="linear", dist_type="pcd",
            data_type="linear", use_threshold=True, data_normalizer=norms[0], data_norm=data_norm)
    m.load_state_dict({k: torch.tensor(z["var:" + k], dtype=torch.float32) for k in vs.get_collection(m.name)})
    ds = I.SemiDataSet(dd, input_size=F, seed=633)
    rep = dist_eval(None, m, B, ds)
    assert rep.accuracy == pytest.approx(float(z["eval_accuracy"]), abs=1e-12)
    assert rep.error == pytest.approx(float(z["eval_error"]), abs=1e-12)
    assert rep.auc == pytest.approx(float(z["eval_auc"]), abs=1e-12)
    dist_predict(None, m, ds, B, os.path.join(dd, "pred"), "predict.txt")
    got = open(os.path.join(dd, "pred", "predict.txt")).read().splitlines()
    want = str(z["predict_txt"]).splitlines()
    assert len(got) == len(want) == len(z["pos"]) + len(z["neg"])
    for g, w in zip(got, want):
        ga, gm, gb, gs = g.split()
        wa, wm, wb, ws = w.split()
        assert (ga, gm, gb) == (wa, wm, wb)
        assert float(gs) == pytest.approx(float(ws), rel=2e-4, abs=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cfl_pcd", "cfl_monomer_reg", "dist", "siamese_margin"])
def test_graph_train_step_equals_eager(kind):
    """train_step_graph (device-only step replayed as a CUDA graph) follows train_step exactly."""
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.models.dist import Dist
    from cfl.ops import normalizer
    F, d, K, B = 48, 8, 3, 32

    def make():
        vs.reset_default_graph()
        vs.set_seed(5)
        if kind == "dist":
            return Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=B, lr=1e-2, beta1=0.9,
                        beta2=0.999, normalize_value=2.0, data_normalizer=normalizer(2.0, 0.0), reg_const=0.01)
        kw = dict(input_shape=(F,), batch_size=B, latent_size=d, num_components=K, model_type="linear",
                  data_type="linear", lr=1e-2)
        if kind == "cfl_pcd":
            return CFL(dist_type="pcd", use_threshold=True, pos_weight=0.25, **kw)
        if kind == "cfl_monomer_reg":
            return CFL(dist_type="monomer", act_type="tanh", use_threshold=True, reg_const=0.01, **kw)
        return CFL(dist_type="siamese", use_threshold=False, caffe_margin=2.0, pos_weight=0.5, **{**kw, "num_components": 1})

    g = torch.Generator().manual_seed(9)
    batches = [[torch.randn(B, F, generator=g).clamp_(min=0).cuda() for _ in range(4)] for _ in range(7)]
    vals = [[torch.randn(B, F, generator=g).clamp_(min=0).cuda() for _ in range(4)] for _ in range(7)]
    a = make()
    outs = [a.train_step(*b, val_batches=v) for b, v in zip(batches, vals)]
    pa = {k: v.detach().clone() for k, v in vs.get_collection(a.name).items()}
    avg_a = (a.s_accuracy_avg, a.val_s_accuracy_avg, a.s_margin_adapt_avg)
    b_ = make()
    for bt, v in zip(batches, vals):
        b_.train_step_graph(*bt, val_batches=v)
    got = b_.fetch_scalars()
    assert b_._graph is not None and b_._step == 7
    for k, v in vs.get_collection(b_.name).items():
        torch.testing.assert_close(v.detach(), pa[k], rtol=2e-6, atol=1e-7, msg=k)
    for k in ("s_total_loss", "s_thres_loss", "s_accuracy", "s_margins", "s_loss_reg", "val_s_accuracy"):
        assert got[k] == pytest.approx(outs[-1][k], rel=1e-6, abs=1e-9), k
    assert (b_.s_accuracy_avg, b_.val_s_accuracy_avg, b_.s_margin_adapt_avg) == pytest.approx(avg_a, rel=1e-9)
