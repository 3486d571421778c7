"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE (read from /root/reference at
generation time only) -- run here, in the build container, as

    python tests/golden/make_reference_golden.py

* ``ref_cfl_*.npz`` / ``ref_dist_*.npz``: the reference's ``cfl.models.cfl.CFL`` and
  ``cfl.models.dist.Dist`` classes are instantiated with fixed batches and preset variables under the
  eager TensorFlow-API stand-in of tf_shim.py (TensorFlow 1.x is not installable here); every
  distance, score, loss term, accuracy and the gradient of each optimiser's loss w.r.t. its var_list
  is recorded, by the reference's variable names.
* ``ref_dataset.npz``: the reference's ``cfl.input_data.SemiDataSet`` (numpy only) run on a small
  features.b directory -- the exact batch sequences of next_labeled_batch / whole_*_batches.
* ``ref_evaluate_total.npz``: the reference's ``cfl.bin.evaluate_total`` (numpy + sklearn) report
  lines on synthetic predict files.

The files written are small and committed; tests never import the reference.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, REFERENCE)
from cfl import ops as ref_ops  # noqa: E402
from cfl.models import cfl as ref_cfl  # noqa: E402
from cfl.models import dist as ref_dist  # noqa: E402

assert ref_cfl.__file__.startswith(REFERENCE)


def _np(t):
    return np.array(t.v.detach().numpy(), dtype=np.float64)


def _batches(rng, B, F, scale):
    mk = lambda: np.maximum(rng.normal(size=(B, F)), 0.0) * scale
    return [mk() for _ in range(4)]


def _wn_head(rng, prefix, F, N, bias):
    out = {prefix + "/fully_connected/g": rng.uniform(0.5, 1.5, N),
           prefix + "/fully_connected/V": rng.uniform(-1, 1, (F, N)) * (6.0 / (F + N)) ** 0.5}
    if bias:
        out[prefix + "/fully_connected/biases"] = rng.normal(size=N) * 0.1
    return out


def cfl_case(tag, *, F, d, K, B, dist_type, act_type=None, pos_weight=None, use_threshold=True, caffe_margin=None,
             lambda_m=0.0, reg_const=0.0, directed=False, data_norm=None, theta=0.7, seed=0):
    rng = np.random.default_rng(seed)
    tf_shim.reset_default_graph()
    scale = (data_norm[0] if data_norm else 1.0) * 0.5
    train, val, unl = _batches(rng, B, F, scale), _batches(rng, B, F, scale), _batches(rng, B, F, scale)
    presets = {"CFL/Thresholder/threshold/threshold": np.float64(theta)}
    bias = dist_type.startswith("pcd")
    for enc in (("DistEncoderSrc", "DistEncoderDst") if directed else ("DistEncoder",)):
        presets.update(_wn_head(rng, "CFL/%s/outputs" % enc, F, d, bias))
        if dist_type in ("pcd", "monomer"):
            presets.update(_wn_head(rng, "CFL/%s/prototype_outputs" % enc, F, d * K, bias))
        if dist_type == "monomer":
            presets.update(_wn_head(rng, "CFL/%s/monomer_outputs" % enc, d, K, False))
    tf_shim.preset_variables(presets)
    norms = ref_ops.dist_normalizer(input_shape=(F,), ae_shape=None, data_scale=None, data_mean=None,
                                    data_norm=data_norm, latent_norm=None, data_type="linear")
    as_t = lambda bs: [tf_shim.T(b) for b in bs]
    model = ref_cfl.CFL(
        is_double=False, disable_double=False, latent_shape=None, source_shape=None, input_shape=(F,), ae_shape=None,
        batch_size=B, data_norm=data_norm, data_type="linear", num_components=K, pos_weight=pos_weight, latent_size=d,
        caffe_margin=caffe_margin, gan=False, cgan=False, t_dim=None, dist_type=dist_type, act_type=act_type,
        use_threshold=use_threshold, lr=1e-3, beta1=0.9, beta2=0.999, z_dim=20, z_stddev=1.0, g_dim=64, g_lr=2e-4,
        g_beta1=0.5, g_beta2=0.999, m_prj=None, m_enc=None, d_dim=64, d_lr=2e-4, d_beta1=0.5, d_beta2=0.999,
        lambda_gp=None, lambda_m=lambda_m, lambda_dra=0.5, directed=directed, data_directed=False, model_type="linear",
        gan_type="conv", reg_const=reg_const, batches=as_t(train), val_batches=as_t(val), unlabeled_batches=as_t(unl),
        data_normalizer=norms[0], data_unnormalizer=norms[1], ae_normalizer=norms[2], ae_unnormalizer=norms[3],
        latent_normalizer=norms[4], run_tag=None)
    unused = set(presets) - set(tf_shim.STATE.created)
    assert not unused, "preset names the reference never created: %s" % sorted(unused)
    assert set(tf_shim.STATE.created) == set(presets), sorted(set(tf_shim.STATE.created) - set(presets))
    out = {"meta_name": np.array(model.get_name()),
           "meta_variables": np.array(tf_shim.STATE.created),
           "meta_s_vars": np.array([v.name[:-2] for v in model.s_vars]),
           "meta_th_vars": np.array([v.name[:-2] for v in model.th_vars]),
           "cfg": np.array(repr(dict(F=F, d=d, K=K, B=B, dist_type=dist_type, act_type=act_type, pos_weight=pos_weight,
                                     use_threshold=use_threshold, caffe_margin=caffe_margin, lambda_m=lambda_m,
                                     reg_const=reg_const, directed=directed, data_norm=data_norm)))}
    for i, nm in enumerate(("pos_source", "pos_target", "neg_source", "neg_target")):
        out["in_" + nm], out["in_val_" + nm] = train[i], val[i]
    for k, v in presets.items():
        out["var:" + k] = v
    for nm in ("s_pos_dists", "s_neg_dists", "val_s_pos_dists", "val_s_neg_dists", "s_p_loss_pos", "s_p_loss_neg",
               "s_thres_loss", "s_total_loss", "s_loss_reg", "s_margins", "s_pos_dists_adapt", "s_neg_dists_adapt",
               "s_margin_adapt", "s_accuracy", "val_s_accuracy"):
        out["out_" + nm] = _np(getattr(model, nm))
    if caffe_margin or lambda_m:
        out["out_s_cd_loss"] = _np(model.s_cd_loss)
    for nm in ("s_pos_predicts", "s_neg_predicts", "val_s_pos_predicts", "val_s_neg_predicts"):
        out["out_" + nm] = _np(getattr(model, nm).outputs)
    out["out_threshold"] = _np(model.s_pos_predicts.threshold)
    out["out_src_activations"] = _np(model.s_pos_src.activations)
    if dist_type in ("pcd", "monomer"):
        out["out_src_prototype_activations"] = _np(model.s_pos_src.prototype_activations)
    for k, g in model.s_optim.gradients().items():
        out["grad_s_optim:" + k] = g
    if not use_threshold:
        for k, g in model.th_optim.gradients().items():
            out["grad_th_optim:" + k] = g
    np.savez_compressed(os.path.join(HERE, "ref_cfl_%s.npz" % tag), **out)
    print("ref_cfl_%s: %s  total=%.6f  vars=%d" % (tag, model.get_name(), out["out_s_total_loss"], len(presets)))


def dist_case(tag, *, F, d, K, B, normalize_value, reg_const=0.0, theta=0.4, seed=0):
    rng = np.random.default_rng(seed)
    tf_shim.reset_default_graph()
    train, val = _batches(rng, B, F, normalize_value * 0.5), _batches(rng, B, F, normalize_value * 0.5)
    presets = {"Dist/Thresholder/threshold/threshold": np.float64(theta)}
    for head, n in (("latent_outputs", d), ("pcd_outputs", d * K)):
        presets["Dist/Encoder/%s/fully_connected/weights" % head] = rng.uniform(-1, 1, (F, n)) * (6.0 / (F + n)) ** 0.5
        presets["Dist/Encoder/%s/fully_connected/biases" % head] = rng.normal(size=n) * 0.1
    tf_shim.preset_variables(presets)
    as_t = lambda bs: [tf_shim.T(b) for b in bs]
    model = ref_dist.Dist(input_shape=(F,), latent_size=d, num_components=K, batch_size=B, lr=1e-3, beta1=0.9,
                          beta2=0.999, batches=as_t(train), val_batches=as_t(val), normalize_value=normalize_value,
                          data_normalizer=ref_ops.normalizer(normalize_value, 0.0, None, None),
                          data_unnormalizer=ref_ops.unnormalizer(normalize_value, 0.0), reg_const=reg_const)
    assert set(tf_shim.STATE.created) == set(presets), (sorted(tf_shim.STATE.created), sorted(presets))
    out = {"meta_name": np.array(model.get_name()), "meta_variables": np.array(tf_shim.STATE.created),
           "cfg": np.array(repr(dict(F=F, d=d, K=K, B=B, normalize_value=normalize_value, reg_const=reg_const)))}
    for i, nm in enumerate(("pos_source", "pos_target", "neg_source", "neg_target")):
        out["in_" + nm], out["in_val_" + nm] = train[i], val[i]
    for k, v in presets.items():
        out["var:" + k] = v
    for nm in ("s_pos_dists", "s_neg_dists", "val_s_pos_dists", "val_s_neg_dists", "s_p_loss_pos", "s_p_loss_neg",
               "thres_loss", "s_total_loss", "s_margins", "s_accuracy", "val_s_accuracy"):
        out["out_" + nm] = _np(getattr(model, nm))
    for nm in ("s_pos_predicts", "s_neg_predicts", "val_s_pos_predicts", "val_s_neg_predicts"):
        out["out_" + nm] = _np(getattr(model, nm).outputs)
    for k, g in model.s_optim.gradients().items():
        out["grad_s_optim:" + k] = g
    np.savez_compressed(os.path.join(HERE, "ref_dist_%s.npz" % tag), **out)
    print("ref_dist_%s: %s  total=%.6f" % (tag, model.get_name(), out["out_s_total_loss"]))


# ---- dataset layer --------------------------------------------------------------------------------
def dataset_case():
    from cfl import input_data as ref_input
    rng = np.random.default_rng(11)
    n, F, n_pos, n_neg = 29, 5, 19, 33
    ids = ["B%09d" % i for i in range(n)]
    feats = rng.normal(size=(n, F)).astype(np.float32)
    pos, neg = rng.integers(0, n, (n_pos, 2)), rng.integers(0, n, (n_neg, 2))
    out = {"ids": np.array(ids), "feats": feats, "pos": pos, "neg": neg}
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "features.b"), "wb") as f:
            for i, x in zip(ids, feats):
                f.write(i.encode("ascii"))
                f.write(x.astype("<f4").tobytes())
        for name, pairs in (("pairs_pos.txt", pos), ("pairs_neg.txt", neg)):
            with open(os.path.join(d, name), "w") as f:
                for a, b in pairs:
                    f.write("%s match %s\n" % (ids[a], ids[b]))
        open(os.path.join(d, "source.txt"), "w").write("\n".join(ids[:11]) + "\n")
        open(os.path.join(d, "target.txt"), "w").write("\n".join(ids[11:]) + "\n")
        out["features_b"] = np.frombuffer(open(os.path.join(d, "features.b"), "rb").read(), dtype=np.uint8)
        for bs, switch in ((4, False), (4, True), (19, True), (25, False), (40, True)):
            ds = ref_input.SemiDataSet(d, input_size=F, data_switch=switch, seed=633)
            seq = [np.stack(ds.next_labeled_batch(bs)) for _ in range(14)]
            out["labeled_bs%d_sw%d" % (bs, int(switch))] = np.stack(seq)          # [steps, 4, bs, F]
        ds = ref_input.SemiDataSet(d, input_size=F, seed=5, directed=True)
        out["num_examples"] = np.array(ds.num_examples)
        out["whole_pos_bs6"] = np.concatenate([np.stack(b[:2], 1) for b in ds.whole_pos_batches(6)])
        out["whole_neg_bs7"] = np.concatenate([np.stack(b[:2], 1) for b in ds.whole_neg_batches(7)])
        out["whole_pos_ids"] = np.array(sum((list(b[2]) for b in ds.whole_pos_batches(6, source_ids=True)), []))
        out["unlabeled_bs5"] = np.stack([ds.next_unlabeled_batch(5)[0] for _ in range(9)])
        out["source_indices"], out["target_indices"] = ds.source_indices, ds.target_indices
        out["by_positions"] = ref_input.load_features_by_positions(os.path.join(d, "features.b"), [3, 0, 28], F)
        out["asins_by_positions"] = np.array(ref_input.load_asins_by_positions(os.path.join(d, "features.b"), [3, 0, 28], F))
    np.savez_compressed(os.path.join(HERE, "ref_dataset.npz"), **out)
    print("ref_dataset: num_examples=%d" % out["num_examples"])


# ---- evaluate_total -------------------------------------------------------------------------------
def evaluate_total_case():
    from cfl.bin import evaluate_total as ref_eval
    rng = np.random.default_rng(3)
    ids = ["I%09d" % i for i in range(30)]
    files, lines = {}, {}
    with tempfile.TemporaryDirectory() as root:
        data = os.path.join(root, "data")
        truth = {}
        for split in ("train", "val", "test"):
            os.makedirs(os.path.join(data, split))
            prs = rng.permutation(900)[:70]
            pairs = [(ids[p // 30], ids[p % 30]) for p in prs]
            truth[split] = (pairs[:30], pairs[30:])
            for name, ps in (("pairs_pos.txt", pairs[:30]), ("pairs_neg.txt", pairs[30:])):
                txt = "".join("%s match %s\n" % p for p in ps)
                files["data/%s/%s" % (split, name)] = txt
                open(os.path.join(data, split, name), "w").write(txt)
        preds = []
        for run in range(3):
            pd_ = os.path.join(root, "pred%d" % run)
            os.makedirs(pd_)
            preds.append(pd_)
            for suffix in ("", "_acc"):
                for split, name in (("train", "predict_train%s.txt"), ("val", "predict_val%s.txt"), ("test", "predict%s.txt")):
                    if run == 2 and split == "train":
                        continue                      # a run without train predictions -> -1 columns
                    pos, neg = truth[split]
                    sc = np.r_[rng.normal(0.4 + 0.2 * run, 1, len(pos)), rng.normal(-0.4, 1, len(neg))]
                    if run == 1:
                        sc = np.round(sc * 2) / 2     # ties
                    txt = "".join("%s match %s %s\n" % (a, b, np.float32(s)) for (a, b), s in zip(pos + neg, sc))
                    files["pred%d/%s" % (run, name % suffix)] = txt
                    open(os.path.join(pd_, name % suffix), "w").write(txt)
        def run_ref(**kw):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ref_eval.evaluate(**kw)
            return buf.getvalue()
        base = dict(data_path=[data], only_larger=None)
        lines["best_acc"] = run_ref(predict_paths=preds, select_auc=False, name="m", avg=False, auc_model=False, **base)
        lines["best_auc"] = run_ref(predict_paths=preds, select_auc=True, name="m", avg=False, auc_model=True, **base)
        lines["avg_auc_model"] = run_ref(predict_paths=preds[:2], select_auc=False, name="avg", avg=True, auc_model=True, **base)
        lines["avg_with_missing_train"] = run_ref(predict_paths=preds, select_auc=False, name="avg3", avg=True, auc_model=False, **base)
        lines["only_larger_1"] = run_ref(data_path=[data], predict_paths=preds[:1], select_auc=True, name="ol", avg=False,
                                         auc_model=True, only_larger=1)
        lines["two_data_paths"] = run_ref(data_path=[data, data], predict_paths=preds[:2], select_auc=True, name="two",
                                          avg=False, auc_model=True, only_larger=None)
    np.savez_compressed(os.path.join(HERE, "ref_evaluate_total.npz"),
                        file_names=np.array(list(files)), file_texts=np.array(list(files.values())),
                        line_names=np.array(list(lines)), line_texts=np.array(list(lines.values())))
    for k, v in lines.items():
        print("ref_evaluate_total[%s]: %s" % (k, v.strip()))


# ---- dist_eval / dist_predict over a dataset directory -----------------------------------------------
class _EagerSession(object):
    """``sess.run(model.val_s_pos_predicts.outputs, feed_dict)`` for an eager stand-in: the fed
    batches go through the reference's own val branch (cfl/models/cfl.py:710-722: dist_fn with
    reuse=True on source and target, build_dist, thres_fn with reuse=True)."""

    def __init__(self, model):
        self.model = model

    def run(self, fetch, feed_dict=None):
        m = self.model
        assert fetch is m.val_s_pos_predicts.outputs
        src = tf_shim.T(np.asarray(feed_dict[m.val_input_data_pos_source], dtype=np.float64))
        dst = tf_shim.T(np.asarray(feed_dict[m.val_input_data_pos_target], dtype=np.float64))
        with tf.variable_scope(m.scope, reuse=True):
            s_ = m.dist_fn(m.data_normalizer(src), name=m.dist_name_src, reuse=True)
            t_ = m.dist_fn(m.data_normalizer(dst), name=m.dist_name_dst, reuse=True)
            return _np(m.thres_fn(s_.build_dist(t_), reuse=True).outputs)


def eval_predict_case():
    from cfl import input_data as ref_input
    from cfl import utils as ref_utils
    rng = np.random.default_rng(21)
    n, F, d, K, B = 60, 12, 6, 3, 16
    ids = ["B%09d" % i for i in range(n)]
    feats = (np.maximum(rng.normal(size=(n, F)), 0) * 4).astype(np.float32)
    pos, neg = rng.integers(0, n, (37, 2)), rng.integers(0, n, (53, 2))
    tf_shim.reset_default_graph()
    presets = {"CFL/Thresholder/threshold/threshold": np.float64(1.3)}
    presets.update(_wn_head(rng, "CFL/DistEncoder/outputs", F, d, True))
    presets.update(_wn_head(rng, "CFL/DistEncoder/prototype_outputs", F, d * K, True))
    tf_shim.preset_variables(presets)
    data_norm = (4.0,)
    norms = ref_ops.dist_normalizer(input_shape=(F,), ae_shape=None, data_scale=None, data_mean=None,
                                    data_norm=data_norm, latent_norm=None, data_type="linear")
    dummy = [tf_shim.T(np.zeros((B, F))) for _ in range(4)]
    model = ref_cfl.CFL(
        is_double=False, disable_double=False, latent_shape=None, source_shape=None, input_shape=(F,), ae_shape=None,
        batch_size=B, data_norm=data_norm, data_type="linear", num_components=K, pos_weight=None, latent_size=d,
        caffe_margin=None, gan=False, cgan=False, t_dim=None, dist_type="pcd", act_type=None, use_threshold=True,
        lr=1e-3, beta1=0.9, beta2=0.999, z_dim=20, z_stddev=1.0, g_dim=64, g_lr=2e-4, g_beta1=0.5, g_beta2=0.999,
        m_prj=None, m_enc=None, d_dim=64, d_lr=2e-4, d_beta1=0.5, d_beta2=0.999, lambda_gp=None, lambda_m=0.0,
        lambda_dra=0.5, directed=False, data_directed=False, model_type="linear", gan_type="conv", reg_const=0.0,
        batches=dummy, val_batches=list(dummy), unlabeled_batches=list(dummy), data_normalizer=norms[0],
        data_unnormalizer=norms[1], ae_normalizer=norms[2], ae_unnormalizer=norms[3], latent_normalizer=norms[4])
    out = {"ids": np.array(ids), "feats": feats, "pos": pos, "neg": neg, "F": F, "d": d, "K": K, "B": B,
           "data_norm": np.array(data_norm)}
    for k, v in presets.items():
        out["var:" + k] = v
    with tempfile.TemporaryDirectory() as dd:
        with open(os.path.join(dd, "features.b"), "wb") as f:
            for i, x in zip(ids, feats):
                f.write(i.encode("ascii"))
                f.write(x.astype("<f4").tobytes())
        for name, pairs in (("pairs_pos.txt", pos), ("pairs_neg.txt", neg)):
            with open(os.path.join(dd, name), "w") as f:
                for a, b in pairs:
                    f.write("%s match %s\n" % (ids[a], ids[b]))
        ds = ref_input.SemiDataSet(dd, input_size=F, seed=633)
        sess = _EagerSession(model)
        rep = ref_utils.dist_eval(sess, model, B, ds)
        out["eval_accuracy"], out["eval_error"], out["eval_auc"] = rep.accuracy, rep.error, rep.auc
        ref_utils.dist_predict(sess, model, ds, B, os.path.join(dd, "pred"), "predict.txt")
        out["predict_txt"] = np.array(open(os.path.join(dd, "pred", "predict.txt")).read())
    np.savez_compressed(os.path.join(HERE, "ref_eval_predict.npz"), **out)
    print("ref_eval_predict: accuracy=%.4f auc=%.6f, %d lines" % (rep.accuracy, rep.auc, len(str(out["predict_txt"]).splitlines())))


if __name__ == "__main__":
    cfl_case("pcd_k3_dyadic", F=48, d=16, K=3, B=12, dist_type="pcd", pos_weight=0.0625, data_norm=(31.9098,), seed=1)
    cfl_case("pcd_k1", F=20, d=8, K=1, B=9, dist_type="pcd", seed=2)
    cfl_case("pcd_k2_tanh_reg", F=24, d=8, K=2, B=10, dist_type="pcd", act_type="tanh", reg_const=0.01, seed=3)
    cfl_case("pcd_k4_lambda_m", F=24, d=6, K=4, B=10, dist_type="pcd", lambda_m=0.3, pos_weight=0.5, seed=4)
    cfl_case("pcd_k3_directed_relu", F=24, d=8, K=3, B=10, dist_type="pcd", act_type="relu", directed=True,
             reg_const=0.002, seed=5)
    cfl_case("pcd_k2_theta_at_floor", F=16, d=6, K=2, B=8, dist_type="pcd", theta=1e-6, seed=6)
    cfl_case("monomer_k3_relu", F=24, d=8, K=3, B=10, dist_type="monomer", act_type="relu", pos_weight=0.25, seed=7)
    cfl_case("monomer_k3_reg", F=20, d=6, K=3, B=9, dist_type="monomer", act_type="tanh", reg_const=0.02, seed=13)
    cfl_case("monomer_k2_directed", F=20, d=6, K=2, B=8, dist_type="monomer", directed=True, seed=8)
    cfl_case("siamese_caffe_margin", F=24, d=12, K=1, B=10, dist_type="siamese", caffe_margin=2.0, pos_weight=0.0625,
             use_threshold=False, seed=9)
    cfl_case("siamese_ut_sigmoid", F=24, d=12, K=1, B=10, dist_type="siamese", act_type="sigmoid", seed=10)
    dist_case("k4_monomer_amazon", F=40, d=10, K=4, B=12, normalize_value=58.388599, seed=11)
    dist_case("k1_reg", F=20, d=6, K=1, B=8, normalize_value=2.0, reg_const=0.05, seed=12)
    dataset_case()
    evaluate_total_case()
    eval_predict_case()
