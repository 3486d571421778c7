"""GPU end-to-end over the reference's file formats: features.b + pair files on
disk -> cfl.bin.train -> checkpoints/best stats -> cfl.bin.predict -> cfl.bin.evaluate_total, with the
numbers cross-checked between the device path (dist_eval / cfl_auc) and the file path."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_dataset(root, F=32, C=6, n_items=300, n_pairs=1500, seed=0):
    """Items in C clusters; cluster c is compatible with c+1 and c+3 (two modes -> needs K=2)."""
    from cfl import input_data as I
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(C, F)).astype(np.float32) * 3
    for si, split in enumerate(("train", "val", "test")):
        d = os.path.join(root, split)
        os.makedirs(d, exist_ok=True)
        lab = rng.integers(0, C, n_items)
        X = np.maximum(cent[lab] + rng.normal(size=(n_items, F)).astype(np.float32), 0)
        ids = ["%s%09d" % ("TVX"[si], i) for i in range(n_items)]
        I.write_features(os.path.join(d, "features.b"), ids, X)
        pos, neg = set(), set()
        while len(pos) < n_pairs or len(neg) < n_pairs:
            a, b = rng.integers(0, n_items, 2)
            ok = (lab[b] - lab[a]) % C in (1, 3)
            (pos if ok else neg).add((a, b))
        for name, pairs in (("pairs_pos.txt", sorted(pos)[:n_pairs]), ("pairs_neg.txt", sorted(neg)[:n_pairs])):
            with open(os.path.join(d, name), "w") as f:
                for a, b in pairs:
                    f.write(f"{ids[a]} match {ids[b]}\n")


def test_cfl_train_predict_evaluate_roundtrip(tmp_path, capsys):
    from cfl import variables as vs
    from cfl.bin import evaluate_total, predict, train
    from cfl.utils import Session, dist_eval, get_checkpoint_state, load_best_stats, load_model
    vs.reset_default_graph()
    root = str(tmp_path)
    _make_dataset(os.path.join(root, "parsed_data", "toy"))
    common = ["--data-name", "toy", "--data-root", os.path.join(root, "parsed_data"), "--checkpoint-root",
              os.path.join(root, "checkpoints"), "--log-root", os.path.join(root, "logs"), "--model-type", "linear",
              "--data-type", "linear", "--data-norm", "8.0", "--input-shape", "32", "--dist-type", "pcd",
              "--use-threshold", "--num-components", "2", "--latent-size", "16", "--lr", "0.03"]
    model = train.main(common + ["--batch-size", "100", "--epochs", "6", "--data-switch"])
    name = "cfl_pcd_linear_linear_ls_16_nc_2_ut_norm_8.0"
    assert model.get_name() == name
    ck = os.path.join(root, "checkpoints", "toy", name)
    st = get_checkpoint_state(ck)
    assert os.path.basename(st.model_checkpoint_path) == "model-90"        # 6 epochs x 15 batches
    best = load_best_stats(os.path.join(ck, "best_model", "best_accuracy"))
    assert best.best_epoch is not None and best.best_auc > 0.8
    assert os.path.exists(os.path.join(ck, "best_acc_model", "best_accuracy_by_th"))
    assert os.path.exists(os.path.join(root, "logs", "toy", name, "log.log"))

    # resume: nothing left to do at --epochs 6; one more epoch at --epochs 7 continues from iter 91
    vs.reset_default_graph()
    model2 = train.main(common + ["--batch-size", "100", "--epochs", "7", "--data-switch"])
    assert model2._step == 104       # restored at 90; load_model resumes at iter 91 (utils.py:476-477), 14 left
    assert os.path.basename(get_checkpoint_state(ck).model_checkpoint_path) == "model-105"

    vs.reset_default_graph()
    pdir = predict.main(common + ["--batch-size", "500", "--predict-root", os.path.join(root, "predicts")])
    assert pdir == os.path.join(root, "predicts", "toy", name)
    for f in ("predict_train.txt", "predict_val.txt", "predict.txt", "predict_train_acc.txt", "predict_val_acc.txt",
              "predict_acc.txt"):
        assert os.path.exists(os.path.join(pdir, f))
    first = open(os.path.join(pdir, "predict.txt")).readline().split()
    assert len(first) == 4 and first[1] == "match" and len(first[0]) == 10
    capsys.readouterr()
    res = evaluate_total.main(["--data-path", os.path.join(root, "parsed_data", "toy"), "--predict-paths", pdir,
                               "--auc-model", "--select-auc", "--name", "pcd-toy"])
    assert capsys.readouterr().out.strip().endswith("pcd-toy")
    assert res["test"]["auc"] > 0.8 and res["val"]["auc"] == pytest.approx(load_best_stats(
        os.path.join(ck, "best_model", "best_accuracy")).best_auc, abs=1e-12)

    # the file path and the device path agree exactly: same scores, integer AUC counts
    vs.reset_default_graph()
    from cfl.bin._common import build_cfl
    args = predict.parse_args(common + ["--batch-size", "500"])
    data, m, _ = build_cfl(args)
    load_model(Session(m), os.path.join(ck, "best_model"))
    for split, ds in (("val", data.val), ("test", data.test)):
        dev = dist_eval(None, m, 500, ds)
        assert dev.auc == pytest.approx(res[split]["auc"], abs=1e-12)
        assert dev.error == pytest.approx(res[split]["error"], abs=1e-12)


def test_monomer_dist_train_and_predict(tmp_path, capsys):
    from cfl import variables as vs
    from cfl.bin import evaluate_total, predict_dist, train_dist
    vs.reset_default_graph()
    root = str(tmp_path)
    _make_dataset(os.path.join(root, "parsed_data", "mono"), seed=3)
    common = ["--data-name", "mono", "--data-root", os.path.join(root, "parsed_data"), "--checkpoint-root",
              os.path.join(root, "checkpoints"), "--log-root", os.path.join(root, "logs"), "--input-shape", "32",
              "--normalize-value", "8.0", "--num-components", "2", "--latent-size", "16", "--lr", "0.03"]
    model = train_dist.main(common + ["--epochs", "6"])
    assert model.get_name() == "linear_dist_ls_16_nc_2_reg_0.0_norm_8.0"
    vs.reset_default_graph()
    pdir = predict_dist.main(common + ["--predict-root", os.path.join(root, "predicts")])
    res = evaluate_total.main(["--data-path", os.path.join(root, "parsed_data", "mono"), "--predict-paths", pdir,
                               "--name", "mono"])
    assert res["test"]["auc"] > 0.8 and res["train"]["auc"] > 0.8


def test_dataset_gathers_on_device_match_file_reads(tmp_path):
    from cfl import input_data as I
    _make_dataset(str(tmp_path), n_items=200, n_pairs=300)
    ds = I.SemiDataSet(os.path.join(str(tmp_path), "train"), input_size=32, seed=5)
    assert ds.features.is_cuda
    path = os.path.join(str(tmp_path), "train", "features.b")
    sp, dp, sn, dn = ds.next_batch(64)
    np.testing.assert_array_equal(sp.cpu().numpy(), I.load_features_by_positions(path, ds.pairs_pos[:64, 0], 32))
    np.testing.assert_array_equal(dn.cpu().numpy(), I.load_features_by_positions(path, ds.pairs_neg[:64, 1], 32))


def test_tf_checkpoint_round_trip_through_a_model(tmp_path):
    """Variables + Adam slots + step leave as a TF V2 checkpoint under the reference's names and come
    back into a fresh model: same scores, and the next train step continues identically."""
    from cfl import tf_checkpoint as T
    from cfl import variables as vs
    from cfl.models.cfl import CFL
    from cfl.utils import Session, export_tf_checkpoint, load_model
    g = torch.Generator().manual_seed(3)
    F, d, K, B = 40, 8, 3, 64
    mk = lambda: CFL(input_shape=(F,), batch_size=B, latent_size=d, num_components=K, model_type="linear",
                     dist_type="pcd", data_type="linear", use_threshold=True, reg_const=0.001, lr=0.01)
    batch = [torch.randn(B, F, generator=g).clamp_(min=0).cuda() for _ in range(4)]
    vs.reset_default_graph()
    m = mk()
    for _ in range(3):
        m.train_step(*batch)
    ck = str(tmp_path / "ck")
    export_tf_checkpoint(m, os.path.join(ck, "model-2"))
    names = set(T.load_tf_checkpoint(os.path.join(ck, "model-2")))
    assert {"CFL/DistEncoder/outputs/fully_connected/V", "CFL/DistEncoder/prototype_outputs/fully_connected/g",
            "CFL/Thresholder/threshold/threshold", "CFL/DistEncoder/outputs/fully_connected/V/Adam",
            "CFL/DistEncoder/outputs/fully_connected/V/Adam_1", "beta1_power", "beta2_power"} <= names
    want = m.predict(batch[0], batch[1]).clone()
    out_a = m.train_step(*batch)
    after_a = m.predict(batch[0], batch[1]).clone()
    vs.reset_default_graph()
    m2 = mk()
    _, start = load_model(Session(m2), ck)
    assert start == 3 and m2._step == 3
    assert torch.equal(m2.predict(batch[0], batch[1]), want)
    out_b = m2.train_step(*batch)
    assert out_b["s_total_loss"] == out_a["s_total_loss"]
    assert torch.equal(m2.predict(batch[0], batch[1]), after_a)


def test_cfl_train_with_cuda_graph_matches_eager_run(tmp_path):
    """`python -m cfl.bin.train --cuda-graph` writes the same checkpoints and best-model statistics as the
    eager loop on the same data and seed."""
    from cfl import variables as vs
    from cfl.bin import train
    from cfl.utils import load_best_stats
    root = str(tmp_path)
    _make_dataset(os.path.join(root, "parsed_data", "toy"))
    stats = []
    for tag, extra in (("eager", []), ("graph", ["--cuda-graph"])):
        vs.reset_default_graph()
        common = ["--data-name", "toy", "--data-root", os.path.join(root, "parsed_data"), "--checkpoint-root",
                  os.path.join(root, "ck_" + tag), "--log-root", os.path.join(root, "logs_" + tag), "--model-type",
                  "linear", "--data-type", "linear", "--data-norm", "8.0", "--input-shape", "32", "--dist-type", "pcd",
                  "--use-threshold", "--num-components", "2", "--latent-size", "16", "--lr", "0.03",
                  "--batch-size", "100", "--epochs", "4", "--data-switch"]
        m = train.main(common + extra)
        ck = os.path.join(root, "ck_" + tag, "toy", m.get_name())
        stats.append((load_best_stats(os.path.join(ck, "best_model", "best_accuracy")), m._step,
                      {k: v.detach().clone() for k, v in vs.get_collection(m.name).items()}))
    (sa, na, pa), (sb, nb, pb) = stats
    assert na == nb == 60
    assert sa.best_epoch == sb.best_epoch
    assert sb.best_auc == pytest.approx(sa.best_auc, abs=2e-4) and sb.best_accuracy == pytest.approx(sa.best_accuracy, abs=5e-3)
    for k in pa:
        torch.testing.assert_close(pb[k], pa[k], rtol=1e-3, atol=1e-4, msg=k)


def test_rank_host_pipeline_equals_rank():
    """CatalogIndex.rank_host (pinned in / pinned out, copies overlapped with scoring) returns what rank does,
    batch after batch, with the staging slots and output buffers reused."""
    from cfl.ranking import CatalogIndex, EncoderWeights
    g = torch.Generator().manual_seed(11)
    F, K, d, N, Q, k = 64, 3, 16, 30000, 48, 20
    lim = (6.0 / (F + d)) ** 0.5
    w = EncoderWeights(V0=((torch.rand(F, d, generator=g) * 2 - 1) * lim).cuda(),
                       Vp=((torch.rand(F, K * d, generator=g) * 2 - 1) * lim).cuda(), g0=torch.ones(d).cuda(),
                       gp=torch.ones(K * d).cuda(), b0=torch.zeros(d).cuda(), bp=torch.zeros(K * d).cuda(),
                       weight_norm=True, in_scale=0.5)
    from cfl import _native as nat
    X = torch.randn(N, F, generator=g).clamp_(min=0).cuda()
    E, _, _ = nat.project_fwd(X, w.V0, w.g0, w.b0, True, w.in_scale, None)
    index = CatalogIndex(w, E)
    batches = [torch.randn(Q, F, generator=g).clamp_(min=0).pin_memory() for _ in range(5)]
    outs = [(torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int64).pin_memory()) for _ in range(5)]
    events = [index.rank_host(b, k, ov, oi) for b, (ov, oi) in zip(batches, outs)]
    for ev in events:
        ev.synchronize()
    for b, (ov, oi) in zip(batches, outs):
        tv, ti = index.rank(b.cuda(), k)
        assert torch.equal(ti.cpu(), oi) and torch.equal(tv.cpu(), ov)
