"""python -m cfl.bin.evaluate_total -- the report the reference's experiments/*/eval.sh print
(cfl/bin/evaluate_total.py:16-202): per predict directory the error at threshold 0 and the AUC of
train / val / test, then either the run with the best val metric or mean+-std over runs.

Same flags, file names and output line.  Scores are joined to the labelled pairs by (id1, id2) as the
reference does; the metrics are computed on arrays (error by counting, AUC as the exact Mann-Whitney
count 2U/(2 n+ n-) over the sorted scores -- the quantity sklearn's roc_auc_score returns)."""
import argparse
import os
from collections import Counter

import numpy as np

SPLITS = ("train", "val", "test")
MISSING = {"accuracy": -1.0, "error": -1.0, "auc": -1.0}


def load_pairs(path):
    """Lines ``id1 rel id2 [score]`` -> [(id1, id2, score-or-0.0)]."""
    pairs = []
    with open(path) as infile:
        for line in infile:
            tokens = line.split()
            pairs.append((tokens[0], tokens[2], float(tokens[3]) if len(tokens) >= 4 else 0.0))
    return pairs


def load_data_pairs(path, only_larger=None):
    data_pairs = {}
    for split in SPLITS:
        pos = load_pairs(os.path.join(path, split, "pairs_pos.txt"))
        neg = load_pairs(os.path.join(path, split, "pairs_neg.txt"))
        if only_larger:                     # keep sources with more than `only_larger` positive pairs
            counts = Counter(a for a, _, _ in pos)
            pos = [p for p in pos if counts[p[0]] > only_larger]
            neg = [p for p in neg if counts[p[0]] > only_larger]
        data_pairs[split] = {"pos_pairs": pos, "neg_pairs": neg}
    return data_pairs


def exact_auc(y_true, y_score):
    """P(score+ > score-) + P(tie)/2 from ranks with ties averaged."""
    y_true = np.asarray(y_true, dtype=bool)
    y_score = np.asarray(y_score, dtype=np.float64)
    n_pos, n_neg = int(y_true.sum()), int((~y_true).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    order = np.argsort(y_score, kind="mergesort")
    s = y_score[order]
    starts = np.flatnonzero(np.concatenate(([True], s[1:] != s[:-1])))
    ends = np.concatenate((starts[1:], [len(s)]))
    # twice the average 1-based rank of each tie group, kept in integers
    rank2 = np.repeat(starts + ends + 1, ends - starts)
    two_u = int(rank2[y_true[order]].sum()) - n_pos * (n_pos + 1)
    return two_u / (2.0 * n_pos * n_neg)


def evaluate_accuracy_by_th(y_true, y_score, th=0.0):
    y_true, y_score = np.asarray(y_true) > 0, np.asarray(y_score, dtype=np.float64)
    correct = int(((y_score > th) == y_true).sum())
    total = len(y_true)
    return correct / total, (total - correct) / total


def evaluate_accuracy(pos_pairs, neg_pairs, pred_pairs):
    y_true = [1] * len(pos_pairs) + [0] * len(neg_pairs)
    y_score = [pred_pairs[(x, y)] for x, y, _ in pos_pairs] + [pred_pairs[(x, y)] for x, y, _ in neg_pairs]
    accuracy, error = evaluate_accuracy_by_th(y_true, y_score)
    return {"accuracy": accuracy, "error": error, "auc": exact_auc(y_true, y_score), "y_true": y_true,
            "y_score": y_score}


def evaluate_data_set(data_pairs, predict_path, auc_model):
    names = ({"train": "predict_train.txt", "val": "predict_val.txt", "test": "predict.txt"} if auc_model else
             {"train": "predict_train_acc.txt", "val": "predict_val_acc.txt", "test": "predict_acc.txt"})
    results = {}
    for split in SPLITS:
        path = os.path.join(predict_path, names[split])
        if split == "train" and not os.path.exists(path):
            results[split] = dict(MISSING)
            continue
        pred = {(x, y): v for x, y, v in load_pairs(path)}
        results[split] = evaluate_accuracy(data_pairs[split]["pos_pairs"], data_pairs[split]["neg_pairs"], pred)
    return results


def select_best_result(results, select_auc):
    key = "auc" if select_auc else "accuracy"
    best = None
    for result in results:                 # first of equals wins, as in the reference
        if best is None or result["val"][key] > best["val"][key]:
            best = result
    return best


def average_result(results):
    avg = {}
    for split in SPLITS:
        err = np.array([r[split]["error"] for r in results])
        auc = np.array([r[split]["auc"] for r in results])
        avg[split] = {"error": err.mean(), "error_std": err.std(), "auc": auc.mean(), "auc_std": auc.std()}
    return avg


def format_result(result, name, avg):
    if avg:
        cells = ["{:.2%}+-{:.2%}".format(result[s][m], result[s][m + "_std"]) for m in ("error", "auc") for s in SPLITS]
    else:
        cells = ["{:.2%}".format(result[s][m]) for m in ("error", "auc") for s in SPLITS]
    return "\t".join(cells + [name])


def print_result(result, name, avg):
    print(format_result(result, name, avg))


def evaluate(data_path, predict_paths, select_auc, name, avg, auc_model, only_larger):
    if len(data_path) == 1:
        data_pairs = load_data_pairs(data_path[0], only_larger)
        results = [evaluate_data_set(data_pairs, p, auc_model) for p in predict_paths]
    else:
        assert len(data_path) == len(predict_paths)
        results = [evaluate_data_set(load_data_pairs(d, only_larger), p, auc_model)
                   for d, p in zip(data_path, predict_paths)]
    result = average_result(results) if avg else select_best_result(results, select_auc)
    print_result(result, name, avg)
    return result


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--data-path", nargs="+", required=True)
    parser.add_argument("--predict-paths", nargs="+", required=True)
    parser.add_argument("--select-auc", action="store_true")
    parser.add_argument("--avg", action="store_true")
    parser.add_argument("--auc-model", action="store_true")
    parser.add_argument("--only-larger", type=int)
    parser.add_argument("--name", default="model")
    return evaluate(**vars(parser.parse_args(argv)))


if __name__ == "__main__":
    main()
