"""The oracle pinned against fixtures produced by executing the reference's own code
(tests/golden/make_reference_golden.py: cfl.models.cfl.CFL / cfl.models.dist.Dist under the eager
TF-API stand-in tests/golden/tf_shim.py; cfl.input_data.SemiDataSet and cfl.bin.evaluate_total
natively).  CPU only; nothing here touches /root/reference."""
import ast
import glob
import os

import numpy as np
import pytest

from oracle import cfl_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFL_CASES = sorted(glob.glob(os.path.join(GOLDEN, "ref_cfl_*.npz")))
DIST_CASES = sorted(glob.glob(os.path.join(GOLDEN, "ref_dist_*.npz")))
TOL = dict(rtol=1e-11, atol=1e-13)


def load_case(path):
    z = np.load(path)
    cfg = ast.literal_eval(str(z["cfg"]))
    return z, cfg


def encoder_params(z, cfg, enc):
    """(V, g, b) per head of one encoder, from the reference's variable names."""
    out = {}
    for head in ("outputs", "prototype_outputs", "monomer_outputs"):
        k = "var:CFL/%s/%s/fully_connected/" % (enc, head)
        if k + "V" in z:
            out[head] = (z[k + "V"], z[k + "g"], z[k + "biases"] if k + "biases" in z else None)
    return out


def oracle_pair(cfg, ps, pt, xs, xt):
    """Distances of one batch + everything the backward needs."""
    K, d, act, kind = cfg["K"], cfg["d"], cfg["act_type"], cfg["dist_type"]
    scale = 1.0 / cfg["data_norm"][0] if cfg["data_norm"] else 1.0
    xs, xt = xs * scale, xt * scale
    S = O.build_prototypes(xs, ps, kind, K, d, act)
    Tt = O.build_prototypes(xt, pt, kind, K, d, act)
    if kind == "pcd":
        dist = O.pcd_dist(Tt["activations"], S["prototype_activations"])
    elif kind == "monomer":
        dist = O.monomer_dist(S["activations"], Tt["prototype_activations"], S["monomer_activations"])
    else:
        dist = O.siamese_dist(S["activations"], Tt["activations"])
    return dist, S, Tt, xs, xt


def oracle_pair_bwd(cfg, ps, pt, S, Tt, xs, xt, ddist, grads, names):
    """Accumulate d total / d variables with the oracle's hand-derived backward."""
    K, d, act, kind = cfg["K"], cfg["d"], cfg["act_type"], cfg["dist_type"]

    def head_bwd(enc, head, params, x, z_pre, y_act, dy_act):
        V, g, b = params
        dpre = dy_act * O.activation_grad(y_act, z_pre, act) if act else dy_act
        return head_bwd_pre(enc, head, params, x, dpre)

    def head_bwd_pre(enc, head, params, x, dpre):
        V, g, b = params
        dV, dg, db = O.fc_weight_norm_bwd(x, V, g, x @ V, dpre)
        base = "CFL/%s/%s/fully_connected/" % (enc, head)
        grads[base + "V"] = grads.get(base + "V", 0) + dV
        grads[base + "g"] = grads.get(base + "g", 0) + dg
        if b is not None:
            grads[base + "biases"] = grads.get(base + "biases", 0) + db

    es, et = names
    if kind == "pcd":
        dv, dP = O.pcd_dist_bwd(Tt["activations"], S["prototype_activations"], ddist)
        head_bwd(et, "outputs", pt["outputs"], xt, Tt["outputs"], Tt["activations"], dv)
        flat = S["flat_prototype_activations"]
        pre = O.fc_weight_norm(xs, *ps["prototype_outputs"], None)
        head_bwd(es, "prototype_outputs", ps["prototype_outputs"], xs, pre, flat, dP.reshape(flat.shape))
    elif kind == "siamese":
        r = S["activations"] - Tt["activations"]
        head_bwd(es, "outputs", ps["outputs"], xs, S["outputs"], S["activations"], 2 * r * ddist[:, None])
        head_bwd(et, "outputs", pt["outputs"], xt, Tt["outputs"], Tt["activations"], -2 * r * ddist[:, None])
    else:
        da, dPt, dw = O.monomer_dist_bwd(S["activations"], Tt["prototype_activations"], S["monomer_activations"], ddist)
        dlogit = O.softmax_bwd(S["monomer_activations"], dw)
        Vm, gm, _ = ps["monomer_outputs"]
        head_bwd_pre(es, "monomer_outputs", ps["monomer_outputs"], S["outputs"], dlogit)
        dpre_gate = (dlogit * (gm / np.sqrt((Vm * Vm).sum(0)))[None, :]) @ Vm.T          # gate reads pre-act e0
        dpre = da * O.activation_grad(S["activations"], S["outputs"], act) if act else da
        head_bwd_pre(es, "outputs", ps["outputs"], xs, dpre + dpre_gate)
        flat = Tt["flat_prototype_activations"]
        pre = O.fc_weight_norm(xt, *pt["prototype_outputs"], None)
        head_bwd(et, "prototype_outputs", pt["prototype_outputs"], xt, pre, flat, dPt.reshape(flat.shape))


def test_fixture_inventory():
    assert len(CFL_CASES) == 11 and len(DIST_CASES) == 2
    kinds = {ast.literal_eval(str(np.load(p)["cfg"]))["dist_type"] for p in CFL_CASES}
    assert kinds == {"pcd", "monomer", "siamese"}


@pytest.mark.parametrize("path", CFL_CASES, ids=[os.path.basename(p)[8:-4] for p in CFL_CASES])
def test_oracle_matches_reference_cfl_graph(path):
    z, cfg = load_case(path)
    names = ("DistEncoderSrc", "DistEncoderDst") if cfg["directed"] else ("DistEncoder", "DistEncoder")
    ps, pt = encoder_params(z, cfg, names[0]), encoder_params(z, cfg, names[1])
    theta = float(z["var:CFL/Thresholder/threshold/threshold"])
    reg_vars = []
    for enc in dict.fromkeys(names):
        for V, g, b in encoder_params(z, cfg, enc).values():
            reg_vars += [V] + ([b] if b is not None else [])
    reg = O.l2_reg(cfg["reg_const"], *reg_vars) if cfg["reg_const"] else 0.0
    np.testing.assert_allclose(reg, z["out_s_loss_reg"], **TOL)

    fw = {}
    for tag, pre in (("", "in_"), ("val_", "in_val_")):
        for lab in ("pos", "neg"):
            fw[tag + lab] = oracle_pair(cfg, ps, pt, z[pre + lab + "_source"], z[pre + lab + "_target"])
            np.testing.assert_allclose(fw[tag + lab][0], z["out_%ss_%s_dists" % (tag, lab)][:, 0], **TOL)
    np.testing.assert_allclose(fw["pos"][1]["activations"], z["out_src_activations"], **TOL)
    if cfg["dist_type"] != "siamese":
        np.testing.assert_allclose(fw["pos"][1]["prototype_activations"], z["out_src_prototype_activations"], **TOL)
    dp, dn = fw["pos"][0], fw["neg"][0]
    L = O.dist_losses(dp, dn, theta, pos_weight=cfg["pos_weight"], use_threshold=cfg["use_threshold"],
                      caffe_margin=cfg["caffe_margin"], lambda_m=cfg["lambda_m"], reg=reg)
    np.testing.assert_allclose(L["score_pos"], z["out_s_pos_predicts"][:, 0], **TOL)
    np.testing.assert_allclose(L["score_neg"], z["out_s_neg_predicts"][:, 0], **TOL)
    np.testing.assert_allclose(O.theta_plus(np.float64(theta)), z["out_threshold"], **TOL)
    for mine, ref in (("p_loss_pos", "s_p_loss_pos"), ("p_loss_neg", "s_p_loss_neg"), ("thres_loss", "s_thres_loss"),
                      ("total_loss", "s_total_loss"), ("accuracy", "s_accuracy"), ("margins", "s_margins"),
                      ("pos_dists_adapt", "s_pos_dists_adapt"), ("neg_dists_adapt", "s_neg_dists_adapt")):
        np.testing.assert_allclose(L[mine], z["out_" + ref], err_msg=ref, **TOL)
    if cfg["caffe_margin"] or cfg["lambda_m"]:
        np.testing.assert_allclose(L["cd_loss"], z["out_s_cd_loss"], **TOL)
    Lv = O.dist_losses(fw["val_pos"][0], fw["val_neg"][0], theta)
    np.testing.assert_allclose(Lv["accuracy"], z["out_val_s_accuracy"], **TOL)

    # backward: the hand-derived chain against autograd through the reference's graph
    gp, gn, dth = O.dist_losses_bwd(dp, dn, theta, pos_weight=cfg["pos_weight"], use_threshold=cfg["use_threshold"],
                                    caffe_margin=cfg["caffe_margin"], lambda_m=cfg["lambda_m"])
    grads = {}
    for lab, dd in (("pos", gp), ("neg", gn)):
        _, S, Tt, xs, xt = fw[lab]
        oracle_pair_bwd(cfg, ps, pt, S, Tt, xs, xt, dd, grads, names)
    if cfg["reg_const"]:
        for enc in dict.fromkeys(names):
            for head, (V, g, b) in encoder_params(z, cfg, enc).items():
                base = "CFL/%s/%s/fully_connected/" % (enc, head)
                grads[base + "V"] = grads.get(base + "V", 0) + cfg["reg_const"] * V
                if b is not None:
                    grads[base + "biases"] = grads.get(base + "biases", 0) + cfg["reg_const"] * b
    ref_grads = {k[len("grad_s_optim:"):]: z[k] for k in z.files if k.startswith("grad_s_optim:")}
    th_name = "CFL/Thresholder/threshold/threshold"
    if cfg["use_threshold"]:
        assert th_name in ref_grads                       # theta rides in s_optim's var_list (cfl.py:1076-1083)
        np.testing.assert_allclose(dth, ref_grads.pop(th_name), rtol=1e-10, atol=1e-13)
    else:
        assert th_name not in ref_grads                   # ... otherwise it has its own optimiser on s_thres_loss
        _, _, dth_only = O.dist_losses_bwd(dp, dn, theta, pos_weight=cfg["pos_weight"], use_threshold=True)
        np.testing.assert_allclose(dth_only, z["grad_th_optim:" + th_name], rtol=1e-10, atol=1e-13)
    assert set(grads) <= set(ref_grads)
    for k, g in ref_grads.items():
        if k in grads:
            np.testing.assert_allclose(grads[k], g, rtol=1e-9, atol=1e-12, err_msg=k)
        else:       # directed models own heads the loss never reads (e.g. DistEncoderDst/outputs for monomer)
            assert cfg["directed"] and not np.any(g), k


@pytest.mark.parametrize("path", DIST_CASES, ids=[os.path.basename(p)[9:-4] for p in DIST_CASES])
def test_oracle_matches_reference_dist_graph(path):
    z, cfg = load_case(path)
    K, d = cfg["K"], cfg["d"]
    W0, b0 = (z["var:Dist/Encoder/latent_outputs/fully_connected/" + n] for n in ("weights", "biases"))
    Wp, bp = (z["var:Dist/Encoder/pcd_outputs/fully_connected/" + n] for n in ("weights", "biases"))
    theta = float(z["var:Dist/Thresholder/threshold/threshold"])
    sc = 1.0 / cfg["normalize_value"]
    dist = {}
    for tag, pre in (("", "in_"), ("val_", "in_val_")):
        for lab in ("pos", "neg"):
            _, P = O.fcencoder(z[pre + lab + "_source"] * sc, W0, b0, Wp, bp, K, d)
            v, _ = O.fcencoder(z[pre + lab + "_target"] * sc, W0, b0, Wp, bp, K, d)
            dist[tag + lab] = O.pcd_dist(v, P)
            np.testing.assert_allclose(dist[tag + lab], z["out_%ss_%s_dists" % (tag, lab)][:, 0], **TOL)
    reg = O.l2_reg(cfg["reg_const"], W0, b0, Wp, bp) if cfg["reg_const"] else 0.0
    L = O.dist_losses(dist["pos"], dist["neg"], theta, reg=reg)
    for mine, ref in (("p_loss_pos", "s_p_loss_pos"), ("p_loss_neg", "s_p_loss_neg"), ("thres_loss", "thres_loss"),
                      ("total_loss", "s_total_loss"), ("accuracy", "s_accuracy"), ("margins", "s_margins")):
        np.testing.assert_allclose(L[mine], z["out_" + ref], err_msg=ref, **TOL)
    np.testing.assert_allclose(O.dist_losses(dist["val_pos"], dist["val_neg"], theta)["accuracy"],
                               z["out_val_s_accuracy"], **TOL)
    gp, gn, dth = O.dist_losses_bwd(dist["pos"], dist["neg"], theta)
    gW0, gb0, gWp, gbp = 0, 0, 0, 0
    for lab, dd in (("pos", gp), ("neg", gn)):
        xs, xt = z["in_" + lab + "_source"] * sc, z["in_" + lab + "_target"] * sc
        _, P = O.fcencoder(xs, W0, b0, Wp, bp, K, d)
        v, _ = O.fcencoder(xt, W0, b0, Wp, bp, K, d)
        dv, dP = O.pcd_dist_bwd(v, P, dd)
        gW0, gb0 = gW0 + xt.T @ dv, gb0 + dv.sum(0)
        gWp, gbp = gWp + xs.T @ dP.reshape(len(xs), -1), gbp + dP.reshape(len(xs), -1).sum(0)
    c = cfg["reg_const"]
    want = {"Dist/Encoder/latent_outputs/fully_connected/weights": gW0 + c * W0,
            "Dist/Encoder/latent_outputs/fully_connected/biases": gb0 + c * b0,
            "Dist/Encoder/pcd_outputs/fully_connected/weights": gWp + c * Wp,
            "Dist/Encoder/pcd_outputs/fully_connected/biases": gbp + c * bp,
            "Dist/Thresholder/threshold/threshold": dth}
    ref = {k[len("grad_s_optim:"):]: z[k] for k in z.files if k.startswith("grad_s_optim:")}
    assert set(ref) == set(want)
    for k in want:
        np.testing.assert_allclose(want[k], ref[k], rtol=1e-9, atol=1e-12, err_msg=k)


def test_variable_names_and_directory_names_are_the_references():
    """The names our models register are the ones the reference's graph created."""
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.models.dist import Dist
    from cfl.ops import dist_normalizer, normalizer
    import torch
    vs.set_default_device(torch.device("cpu"))
    for path in CFL_CASES:
        z, cfg = load_case(path)
        vs.reset_default_graph()
        norms = dist_normalizer(input_shape=(cfg["F"],), ae_shape=None, data_scale=None, data_mean=None,
                                data_norm=cfg["data_norm"], latent_norm=None, data_type="linear")
        m = CFL(input_shape=(cfg["F"],), batch_size=cfg["B"], latent_size=cfg["d"], num_components=cfg["K"],
                model_type="linear", dist_type=cfg["dist_type"], act_type=cfg["act_type"], data_type="linear",
                use_threshold=cfg["use_threshold"], pos_weight=cfg["pos_weight"], caffe_margin=cfg["caffe_margin"],
                lambda_m=cfg["lambda_m"], reg_const=cfg["reg_const"], directed=cfg["directed"],
                data_normalizer=norms[0], data_norm=cfg["data_norm"])
        assert m.get_name() == str(z["meta_name"])
        assert sorted(vs.all_variables()) == sorted(str(s) for s in z["meta_variables"])
        assert sorted(k for k, v in vs.all_variables().items() if any(v is p for p in m.s_vars)) == \
            sorted(str(s) for s in z["meta_s_vars"])
        for k, v in vs.all_variables().items():
            assert tuple(v.shape) == z["var:" + k].shape, k
    for path in DIST_CASES:
        z, cfg = load_case(path)
        vs.reset_default_graph()
        m = Dist(input_shape=(cfg["F"],), latent_size=cfg["d"], num_components=cfg["K"], batch_size=cfg["B"], lr=1e-3,
                 beta1=0.9, beta2=0.999, normalize_value=cfg["normalize_value"],
                 data_normalizer=normalizer(cfg["normalize_value"], 0.0), reg_const=cfg["reg_const"])
        assert m.get_name() == str(z["meta_name"])
        assert sorted(vs.all_variables()) == sorted(str(s) for s in z["meta_variables"])


def test_dataset_layer_matches_reference_semidataset(tmp_path):
    from cfl import input_data as I
    z = np.load(os.path.join(GOLDEN, "ref_dataset.npz"))
    ids, feats, pos, neg = [str(s) for s in z["ids"]], z["feats"], z["pos"], z["neg"]
    d = str(tmp_path)
    I.write_features(os.path.join(d, "features.b"), ids, feats)
    assert open(os.path.join(d, "features.b"), "rb").read() == z["features_b"].tobytes()     # byte-exact writer
    for name, pairs in (("pairs_pos.txt", pos), ("pairs_neg.txt", neg)):
        with open(os.path.join(d, name), "w") as f:
            for a, b in pairs:
                f.write("%s match %s\n" % (ids[a], ids[b]))
    open(os.path.join(d, "source.txt"), "w").write("\n".join(ids[:11]) + "\n")
    open(os.path.join(d, "target.txt"), "w").write("\n".join(ids[11:]) + "\n")
    F = feats.shape[1]
    for key in (k for k in z.files if k.startswith("labeled_")):
        bs, sw = int(key.split("_")[1][2:]), bool(int(key.split("_")[2][2:]))
        ds = I.SemiDataSet(d, input_size=F, data_switch=sw, seed=633, device="cpu")
        for step, want in enumerate(z[key]):
            got = np.stack([t.numpy() for t in ds.next_labeled_batch(bs)])
            np.testing.assert_array_equal(got, want, err_msg="%s step %d" % (key, step))
    ds = I.SemiDataSet(d, input_size=F, seed=5, directed=True, device="cpu")
    assert ds.num_examples == int(z["num_examples"])
    np.testing.assert_array_equal(np.concatenate([np.stack([t.numpy() for t in b[:2]], 1) for b in ds.whole_pos_batches(6)]),
                                  z["whole_pos_bs6"])
    np.testing.assert_array_equal(np.concatenate([np.stack([t.numpy() for t in b[:2]], 1) for b in ds.whole_neg_batches(7)]),
                                  z["whole_neg_bs7"])
    assert sum((list(b[2]) for b in ds.whole_pos_batches(6, source_ids=True)), []) == [str(s) for s in z["whole_pos_ids"]]
    np.testing.assert_array_equal(np.stack([ds.next_unlabeled_batch(5)[0].numpy() for _ in range(9)]), z["unlabeled_bs5"])
    np.testing.assert_array_equal(ds.source_indices, z["source_indices"])
    np.testing.assert_array_equal(ds.target_indices, z["target_indices"])
    np.testing.assert_array_equal(I.load_features_by_positions(os.path.join(d, "features.b"), [3, 0, 28], F), z["by_positions"])
    assert I.load_asins_by_positions(os.path.join(d, "features.b"), [3, 0, 28], F) == [str(s) for s in z["asins_by_positions"]]


def test_evaluate_total_prints_the_references_lines(tmp_path, capsys):
    from cfl.bin import evaluate_total as E
    z = np.load(os.path.join(GOLDEN, "ref_evaluate_total.npz"))
    for name, text in zip(z["file_names"], z["file_texts"]):
        p = tmp_path / str(name)
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(str(text))
    lines = dict(zip((str(s) for s in z["line_names"]), (str(s) for s in z["line_texts"])))
    data = str(tmp_path / "data")
    preds = [str(tmp_path / ("pred%d" % i)) for i in range(3)]
    runs = {
        "best_acc": ["--data-path", data, "--predict-paths", *preds, "--name", "m"],
        "best_auc": ["--data-path", data, "--predict-paths", *preds, "--name", "m", "--select-auc", "--auc-model"],
        "avg_auc_model": ["--data-path", data, "--predict-paths", *preds[:2], "--name", "avg", "--avg", "--auc-model"],
        "avg_with_missing_train": ["--data-path", data, "--predict-paths", *preds, "--name", "avg3", "--avg"],
        "only_larger_1": ["--data-path", data, "--predict-paths", preds[0], "--name", "ol", "--select-auc", "--auc-model",
                          "--only-larger", "1"],
        "two_data_paths": ["--data-path", data, data, "--predict-paths", *preds[:2], "--name", "two", "--select-auc",
                           "--auc-model"],
    }
    assert set(runs) == set(lines)
    for key, argv in runs.items():
        capsys.readouterr()
        E.main(argv)
        assert capsys.readouterr().out == lines[key], key
