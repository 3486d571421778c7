"""CPU tests of the predict / evaluate file layer: evaluate_total's metrics against
sklearn and a scalar restatement of the reference's counting loop, its report line, and the
checkpoint-index bookkeeping load_model relies on."""
import os

import numpy as np
import pytest
import torch


def _write_split(root, split, pos, neg):
    d = os.path.join(root, split)
    os.makedirs(d, exist_ok=True)
    for name, pairs in (("pairs_pos.txt", pos), ("pairs_neg.txt", neg)):
        with open(os.path.join(d, name), "w") as f:
            for a, b in pairs:
                f.write(f"{a} match {b}\n")


def _case(tmp_path, seed, ties=False):
    rng = np.random.default_rng(seed)
    data_root, pred_root = str(tmp_path / "data"), str(tmp_path / f"pred{seed}")
    os.makedirs(pred_root, exist_ok=True)
    truth = {}
    ids = ["I%09d" % i for i in range(40)]
    names = {"train": "predict_train.txt", "val": "predict_val.txt", "test": "predict.txt"}
    for split in ("train", "val", "test"):
        prs = rng.permutation(40 * 40)[:90]          # distinct (a, b) so the dict join is unambiguous
        pairs = [(ids[p // 40], ids[p % 40]) for p in prs]
        pos, neg = pairs[:35], pairs[35:]
        _write_split(data_root, split, pos, neg)
        sc = np.r_[rng.normal(0.5, 1, 35), rng.normal(-0.5, 1, 55)]
        if ties:
            sc = np.round(sc)
        truth[split] = (np.r_[np.ones(35), np.zeros(55)], sc)
        with open(os.path.join(pred_root, names[split]), "w") as f:
            for (a, b), s in zip(pairs, sc):
                f.write("{} match {} {}\n".format(a, b, np.float32(s)))
        truth[split] = (truth[split][0], np.float32(sc).astype(np.float64))
    return data_root, pred_root, truth


@pytest.mark.parametrize("ties", [False, True])
def test_evaluate_total_matches_sklearn_and_the_counting_loop(tmp_path, ties):
    from sklearn.metrics import roc_auc_score
    from cfl.bin import evaluate_total as E
    data_root, pred_root, truth = _case(tmp_path, 1, ties)
    res = E.evaluate_data_set(E.load_data_pairs(data_root), pred_root, auc_model=True)
    for split, (y, s) in truth.items():
        assert res[split]["auc"] == pytest.approx(roc_auc_score(y, s), abs=1e-15)
        correct = sum((sc > 0) == (lab > 0) for lab, sc in zip(y, s))      # evaluate_total.py:123-139
        assert res[split]["accuracy"] == correct / len(y)
        assert res[split]["error"] == (len(y) - correct) / len(y)


def test_report_line_selection_average_and_missing_train(tmp_path, capsys):
    from cfl.bin import evaluate_total as E
    data_root, p1, _ = _case(tmp_path, 1)
    _, p2, _ = _case(tmp_path, 2)              # same pairs file set is rewritten with seed-2 pairs
    _, p1, _ = _case(tmp_path, 1)              # restore seed-1 pairs; pred2 now misses keys -> use copies of p1
    import shutil
    p3 = str(tmp_path / "pred3")
    shutil.copytree(p1, p3)
    os.remove(os.path.join(p3, "predict_train.txt"))
    r = E.main(["--data-path", data_root, "--predict-paths", p1, p3, "--auc-model", "--select-auc", "--name", "x"])
    line = capsys.readouterr().out.strip().split("\t")
    assert len(line) == 7 and line[-1] == "x" and all(c.endswith("%") for c in line[:6])
    assert r["train"]["auc"] != -1.0            # ties on val -> the first run wins (it has train scores)
    assert line[0] == "{:.2%}".format(r["train"]["error"])
    a = E.main(["--data-path", data_root, "--predict-paths", p1, p3, "--auc-model", "--avg"])
    cells = capsys.readouterr().out.strip().split("\t")
    assert all("+-" in c for c in cells[:6])
    assert a["train"]["error"] == pytest.approx((r["train"]["error"] - 1.0) / 2)     # missing train counts as -1
    assert a["test"]["auc_std"] == 0.0
    # only_larger keeps sources with more than n positives
    kept = E.load_data_pairs(data_root, only_larger=1)["test"]
    from collections import Counter
    c = Counter(a_ for a_, _, _ in E.load_pairs(os.path.join(data_root, "test", "pairs_pos.txt")))
    assert all(c[a_] > 1 for a_, _, _ in kept["pos_pairs"] + kept["neg_pairs"])


class _ToyModel:
    def __init__(self):
        self.w = torch.zeros(3)

    def state_dict(self):
        return {"CFL/w": self.w.clone(), "CFL/w/Adam": self.w * 2, "__step__": torch.tensor(7)}

    def load_state_dict(self, sd):
        self.loaded = dict(sd)
        self.w = sd["CFL/w"].clone()


def test_saver_index_file_and_load_model(tmp_path):
    from cfl.utils import Saver, Session, get_checkpoint_state, load_best_stats, load_model
    m = _ToyModel()
    sess = Session(m)
    ck = str(tmp_path / "ck")
    saver = Saver(max_to_keep=2)
    for step in (10, 20, 30):
        m.w = torch.full((3,), float(step))
        saver.save(sess, os.path.join(ck, "model"), global_step=step)
    st = get_checkpoint_state(ck)
    assert os.path.basename(st.model_checkpoint_path) == "model-30"
    assert [os.path.basename(p) for p in st.all_model_checkpoint_paths] == ["model-20", "model-30"]
    assert not os.path.exists(os.path.join(ck, "model-10.pt"))
    assert open(os.path.join(ck, "checkpoint")).readline() == 'model_checkpoint_path: "model-30"\n'
    m2 = _ToyModel()
    _, start = load_model(Session(m2), ck)
    assert start == 31 and torch.equal(m2.w, torch.full((3,), 30.0)) and "CFL/w/Adam" in m2.loaded
    assert load_model(Session(_ToyModel()), str(tmp_path / "none"))[1] == 0
    # load_pre_weights: trainable variables of <dir>/best_model, step of <dir>
    saver2 = Saver()
    m.w = torch.full((3,), 5.0)
    saver2.save(sess, os.path.join(ck, "best_model", "model"), global_step=3)
    m3 = _ToyModel()
    _, start = load_model(Session(m3), str(tmp_path / "fresh"), load_pre_weights=ck)
    assert start == 31 and torch.equal(m3.w, torch.full((3,), 5.0)) and "CFL/w/Adam" not in m3.loaded
    with pytest.raises(Exception):
        load_model(Session(_ToyModel()), str(tmp_path / "fresh"), load_pre_weights=str(tmp_path / "nothing"))
    p = tmp_path / "best_accuracy"
    p.write_text("None\t0.5\t0.75")
    s = load_best_stats(str(p))
    assert s.best_epoch is None and s.best_accuracy == 0.5 and s.best_auc == 0.75


def test_bin_parsers_accept_the_experiment_scripts_flags():
    """Flag sets of experiments/dyadic/run.sh and experiments/monomer (train_dist)."""
    from cfl.bin import predict, train, train_dist
    a = train.parse_args("--data-name dyadic_latent --model-type linear --data-type linear --data-norm 31.9098 "
                         "--data-switch --input-shape 1024 --dist-type pcd --pos-weight 0.0625 --use-threshold "
                         "--num-components 3 --latent-size 64 --epochs 5".split())
    assert a.data_switch and a.epochs == 5 and a.input_shape == (1024,) and a.data_norm == (31.9098,)
    b = predict.parse_args("--data-name dyadic_latent --model-type linear --data-type linear --data-norm 31.9098 "
                           "--input-shape 1024 --dist-type siamese --caffe-margin 100. --pos-weight 0.0625 "
                           "--num-components 1 --latent-size 256".split())
    assert b.batch_size == 500 and b.predict_root == "predicts" and b.caffe_margin == 100.0
    c = train_dist.parse_args("--data-name monomer/Baby-also_viewed --latent-size 10 --num-components 4 "
                              "--normalize-value 58.388599 --epochs 3".split())
    assert c.input_shape == (4096,) and c.normalize_value == 58.388599
