"""cfl.utils -- evaluation / prediction loops and the flag parsers of the reference
(cfl/utils.py), with the device-side AUC kernel instead of sklearn on the host.  ``sess`` arguments
are kept for signature compatibility and ignored (there is no TF session)."""
from __future__ import annotations

import argparse
import os
import re
from argparse import Namespace

import torch


def log_args(args):
    """cfl/utils.py:20-22: one warning-level log line per flag, sorted by name."""
    import logging
    logger = logging.getLogger(__name__)
    for name, value in sorted(vars(args).items()):
        logger.warning("%s = %r", name, value)


def reduce_product(xs):
    """cfl/utils.py:25-29."""
    s = 1
    for x in xs:
        s *= x
    return s


def load_best_stats(path):
    """cfl/utils.py:32-42: ``"{epoch}\\t{accuracy}\\t{auc}"``."""
    stats = Namespace(best_epoch=None, best_accuracy=0.0, best_auc=0.0)
    if os.path.exists(path):
        with open(path) as infile:
            tokens = infile.read().split("\t")
            best_epoch, best_accuracy = tokens[:2]
            stats.best_epoch = int(best_epoch) if best_epoch.strip() != "None" else None
            stats.best_accuracy = float(best_accuracy)
            if len(tokens) >= 3:
                stats.best_auc = float(tokens[2])
    return stats


class Session(object):
    """Stand-in for ``tf.Session``: there is no graph, so the only thing a "session" carries is the
    model whose variables a Saver reads and writes."""

    def __init__(self, model=None):
        self.model = model

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def get_checkpoint_state(checkpoint_dir):
    """``tf.train.get_checkpoint_state``: parses the text-proto ``checkpoint`` file the Saver keeps."""
    path = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    latest, every = None, []
    with open(path) as infile:
        for line in infile:
            m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths):\s*"(.*)"', line)
            if not m:
                continue
            full = m.group(2) if os.path.isabs(m.group(2)) else os.path.join(checkpoint_dir, m.group(2))
            if m.group(1) == "model_checkpoint_path":
                latest = full
            else:
                every.append(full)
    return Namespace(model_checkpoint_path=latest, all_model_checkpoint_paths=every) if latest else None


class Saver(object):
    """``tf.train.Saver`` for this package: ``save(sess, prefix, global_step)`` writes
    ``{prefix}-{step}.pt`` (variables under their reference names + Adam slots, see
    ``PairModel.state_dict``) and maintains the ``checkpoint`` index file with the same keys TF uses,
    keeping the newest ``max_to_keep`` files."""

    SUFFIX = ".pt"

    def __init__(self, max_to_keep=5):
        self.max_to_keep = max_to_keep
        self._kept = []

    def save(self, sess, save_path, global_step=None):
        prefix = save_path if global_step is None else "{}-{}".format(save_path, global_step)
        d = os.path.dirname(prefix)
        os.makedirs(d, exist_ok=True)
        torch.save({k: v.cpu() for k, v in sess.model.state_dict().items()}, prefix + self.SUFFIX)
        if prefix in self._kept:
            self._kept.remove(prefix)
        self._kept.append(prefix)
        while self.max_to_keep and len(self._kept) > self.max_to_keep:
            old = self._kept.pop(0)
            if os.path.exists(old + self.SUFFIX):
                os.remove(old + self.SUFFIX)
        with open(os.path.join(d, "checkpoint"), "w") as out:
            out.write('model_checkpoint_path: "{}"\n'.format(os.path.basename(prefix)))
            for k in self._kept:
                out.write('all_model_checkpoint_paths: "{}"\n'.format(os.path.basename(k)))
        return prefix

    def restore(self, sess, save_path, trainable_only=False):
        if not os.path.exists(save_path + self.SUFFIX) and os.path.exists(save_path + ".index"):
            # a checkpoint written by the reference's tf.train.Saver (V2 tensor bundle)
            from .tf_checkpoint import load_tf_checkpoint, tf_to_state_dict
            sd = tf_to_state_dict(load_tf_checkpoint(save_path), beta1=getattr(sess.model, "beta1", 0.9),
                                  beta2=getattr(sess.model, "beta2", 0.999))
        else:
            sd = torch.load(save_path + self.SUFFIX, map_location="cpu")
        if trainable_only:          # assign_from_checkpoint_fn(trainable_variables(), ignore_missing_vars=True)
            sd = {k: v for k, v in sd.items() if not (k.endswith("/Adam") or k.endswith("/Adam_1") or k == "__step__")}
        sess.model.load_state_dict(sd)


def export_tf_checkpoint(model, prefix):
    """Write the model's variables (and Adam slots, ``beta1_power`` / ``beta2_power``) as a TensorFlow V2 checkpoint.

    What the reference can do with it (cfl/utils.py:465-497): a ``Dist`` graph (one optimiser, no moving averages)
    restores it with its plain ``tf.train.Saver().restore``.  A ``CFL`` graph also declares the EMA(0.99) shadow
    variables of its statistics (cfl/models/cfl.py:528,903-949 -- their names contain TF's auto-numbered op names, which
    cannot be reproduced without TensorFlow) and, without ``--use-threshold``, a second optimiser's ``beta*_power_1``:
    those are NOT in the export, so a CFL graph must read it through the reference's own pre-weights route,
    ``assign_from_checkpoint_fn(trainable_variables(), ignore_missing_vars=True)`` (utils.py:481-490), or a Saver with a
    restricted ``var_list``.  No TF-written file has been read by, and no export handed to, a real TensorFlow here."""
    from .tf_checkpoint import state_dict_to_tf, write_tf_checkpoint
    write_tf_checkpoint(prefix, state_dict_to_tf(model.state_dict(), getattr(model, "beta1", 0.9),
                                                 getattr(model, "beta2", 0.999)))
    d = os.path.dirname(os.path.abspath(prefix))
    with open(os.path.join(d, "checkpoint"), "w") as out:
        out.write('model_checkpoint_path: "{0}"\nall_model_checkpoint_paths: "{0}"\n'.format(os.path.basename(prefix)))


def _step_of(checkpoint_path):
    return int(re.search(r"(\d+)", os.path.basename(checkpoint_path).split("-")[-1]).group(1))


def load_model(sess, checkpoint_dir, load_pre_weights=None):
    """cfl/utils.py:465-497: restore the newest checkpoint of ``checkpoint_dir`` and return
    ``(saver, start_step)`` with start_step = its global step + 1; with ``load_pre_weights`` and no
    own checkpoint, take the trainable variables of ``<load_pre_weights>/best_model``."""
    saver = Saver()
    ckpt = get_checkpoint_state(checkpoint_dir)
    if ckpt and ckpt.model_checkpoint_path:
        saver.restore(sess, ckpt.model_checkpoint_path)
        return saver, _step_of(ckpt.model_checkpoint_path) + 1
    start_step = 0
    if load_pre_weights:
        best_model = os.path.join(load_pre_weights, "best_model")
        ckpt, ckpt_all = get_checkpoint_state(best_model), get_checkpoint_state(load_pre_weights)
        if not (ckpt and ckpt_all):
            raise Exception("must have best model! %s" % best_model)
        saver.restore(sess, ckpt.model_checkpoint_path, trainable_only=True)
        start_step = _step_of(ckpt_all.model_checkpoint_path) + 1
    return saver, start_step


class IncrementalAverage(object):
    """cfl/utils.py:500-507."""

    def __init__(self):
        self.average = 0.0
        self.count = 0

    def add(self, value):
        self.count += 1
        self.average = (value - self.average) / self.count + self.average


def _pair_batches(batches, model):
    """Feeds of dist_eval (cfl/utils.py:234-243): (source, target) of a batch tuple; for 'double'
    datasets the latent halves (positions 1 and 3)."""
    if getattr(model, "is_double", False):
        return batches[1], batches[3]
    return batches[0], batches[1]


def dist_eval(sess, model, batch_size, data, with_roc=False):
    """cfl/utils.py:227-274.  Positive pairs then negative pairs go through the model's single
    scoring node; accuracy at threshold 0; AUC.  The scores stay on the GPU: AUC and the
    accuracy counts come from the exact integer rank/count kernel (cfl_auc), AUC =
    twoU / (2 n+ n-), within 2 ulp of sklearn's roc_auc_score."""
    from . import _native as nat
    pos, neg = [], []
    for batches in data.whole_pos_batches(batch_size):
        pos.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    for batches in data.whole_neg_batches(batch_size):
        neg.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    dev = pos[0].device if pos else neg[0].device
    pos = torch.cat(pos) if pos else torch.empty(0, device=dev)
    neg = torch.cat(neg) if neg else torch.empty(0, device=dev)
    two_u, n_pos, n_neg, correct = nat.auc_counts(pos, neg).cpu().tolist()
    total = n_pos + n_neg
    auc = two_u / (2.0 * n_pos * n_neg) if n_pos and n_neg else float("nan")
    roc = None
    if with_roc:                      # optional host-side curve, not on the hot path
        from sklearn.metrics import roc_curve
        y = [1] * n_pos + [0] * n_neg
        roc = roc_curve(y, torch.cat([pos, neg]).cpu().numpy())
    return Namespace(error=(total - correct) / total, accuracy=correct / total, auc=auc, roc=roc,
                     two_u=two_u, n_pos=n_pos, n_neg=n_neg)


def sharded_auc_counts(pos, neg, group=None):
    """AUC / accuracy integers of labelled pairs whose SCORES are sharded over ranks:
    the (smaller) positive-score arrays are all-gathered, every rank counts its own negatives against
    all positives with the exact rank/count kernel (cfl_auc), and the integers are summed
    (``all_reduce``) -- bit-identical for any number of ranks.  ``pos`` / ``neg`` = this rank's scores
    (either may be empty).  -> (two_u, n_pos, n_neg, correct) python ints, the same on every rank."""
    from . import _native as nat
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    pos, neg = pos.reshape(-1).float(), neg.reshape(-1).float()
    if world > 1:
        sizes = torch.zeros(world, dtype=torch.int64, device=pos.device)
        sizes[dist.get_rank(group)] = pos.numel()
        dist.all_reduce(sizes, group=group)
        cap = int(sizes.max())
        padded = torch.zeros(cap, dtype=torch.float32, device=pos.device)
        padded[:pos.numel()] = pos
        gathered = torch.empty(world * cap, dtype=torch.float32, device=pos.device)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        pos = torch.cat([gathered[r * cap:r * cap + int(sizes[r])] for r in range(world)])
    pos_correct = int((pos > 0).sum())
    if neg.numel() and pos.numel():
        two_u, _, n_neg, correct = nat.auc_counts(pos, neg).cpu().tolist()
        part = [two_u, n_neg, correct - pos_correct]
    else:
        part = [0, neg.numel(), int((neg <= 0).sum())]
    if world > 1:
        t = torch.tensor(part, dtype=torch.int64, device=pos.device)
        dist.all_reduce(t, group=group)
        part = t.tolist()
    return part[0], pos.numel(), part[1], part[2] + pos_correct


def dist_eval_sharded(sess, model, batch_size, data, rank=None, world=None, group=None):
    """``dist_eval`` with the labelled pairs dealt to the ranks batch by batch (batch b -> rank b mod world):
    each rank scores its share, ``sharded_auc_counts`` combines the integers.  Same Namespace as ``dist_eval``,
    identical on every rank and identical to the single-process result."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank = (dist.get_rank(group) if on else 0) if rank is None else rank
    world = (dist.get_world_size(group) if on else 1) if world is None else world
    pos, neg = [], []
    for b, batches in enumerate(data.whole_pos_batches(batch_size)):
        if b % world == rank:
            pos.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    for b, batches in enumerate(data.whole_neg_batches(batch_size)):
        if b % world == rank:
            neg.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    # a rank that received no batch still takes part in the collectives: its empty tensors must live on the GPU
    from . import variables as vs
    dev = getattr(model, "device", None) or (pos[0].device if pos else neg[0].device if neg else vs.default_device())
    pos = torch.cat(pos) if pos else torch.empty(0, device=dev)
    neg = torch.cat(neg) if neg else torch.empty(0, device=dev)
    two_u, n_pos, n_neg, correct = sharded_auc_counts(pos, neg, group)
    total = n_pos + n_neg
    auc = two_u / (2.0 * n_pos * n_neg) if n_pos and n_neg else float("nan")
    if total == 0:                                           # no labelled pairs at all
        return Namespace(error=float("nan"), accuracy=float("nan"), auc=auc, roc=None, two_u=0, n_pos=0, n_neg=0)
    return Namespace(error=(total - correct) / total, accuracy=correct / total, auc=auc, roc=None,
                     two_u=two_u, n_pos=n_pos, n_neg=n_neg)


def dist_predict(sess, model, data, batch_size, predict_dir, output_name):
    """cfl/utils.py:277-321: writes ``"{id1} match {id2} {score}\\n"``, positives first."""
    pos, neg = [], []
    for batches in data.whole_pos_batches(batch_size):
        pos.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    for batches in data.whole_neg_batches(batch_size):
        neg.append(model.predict(*_pair_batches(batches, model)).reshape(-1))
    rules = [(torch.cat(pos).cpu().numpy() if pos else [], data.pairs_pos),
             (torch.cat(neg).cpu().numpy() if neg else [], data.pairs_neg)]
    os.makedirs(predict_dir, exist_ok=True)
    with open(os.path.join(predict_dir, output_name), "w") as outfile:
        for dists, pairs in rules:
            for pred, (idx1, idx2) in zip(dists, pairs):
                outfile.write("{} match {} {}\n".format(data.index_to_asins[idx1], data.index_to_asins[idx2], pred))


def dist_check_args(args):
    """cfl/utils.py:45-81 (the checks that concern the distance model)."""
    for name in ("input_shape", "latent_shape", "data_mean", "data_norm"):
        v = getattr(args, name, None)
        if v:
            setattr(args, name, tuple(v))
    if args.caffe_margin and args.caffe_margin > 0:
        assert args.dist_type == "siamese", "only use cd loss in siamese"
    if args.caffe_margin and args.caffe_margin < 0:
        assert args.lambda_m > 0
    if args.lambda_m > 0:
        assert not args.caffe_margin
    if args.dist_type != "siamese":
        assert args.use_threshold, "must use entropy loss"


def dist_parser(data_name="mnist", data_root="parsed_data", checkpoint_root="checkpoints", log_root="logs",
                run_tag=None, seed=633, input_shape=(28, 28, 1), batch_size=100, data_scale=None,
                data_mean=None, data_norm=None, latent_norm=None, data_type="sigmoid", model_type="conv",
                dist_type="pcd", act_type=None, use_threshold=False, lr=0.001, beta1=0.9, beta2=0.999,
                num_components=2, latent_size=20, caffe_margin=None, pos_weight=None, lambda_m=0.0,
                reg_const=0.0, latent_shape=None):
    """The distance-model flags of cfl/utils.py:84-224, same names and defaults (GAN flags omitted)."""
    p = argparse.ArgumentParser()
    p.add_argument("--data-name", default=data_name)
    p.add_argument("--data-root", default=data_root)
    p.add_argument("--checkpoint-root", default=checkpoint_root)
    p.add_argument("--log-root", default=log_root)
    p.add_argument("--run-tag", default=run_tag)
    p.add_argument("--seed", type=int, default=seed)
    p.add_argument("--input-shape", nargs="+", type=int, default=input_shape)
    p.add_argument("--latent-shape", nargs="+", type=int, default=latent_shape)
    p.add_argument("--batch-size", type=int, default=batch_size)
    p.add_argument("--data-scale", type=float, default=data_scale)
    p.add_argument("--data-mean", nargs="+", type=float, default=data_mean)
    p.add_argument("--data-norm", nargs="+", type=float, default=data_norm)
    p.add_argument("--latent-norm", type=float, default=latent_norm)
    p.add_argument("--data-type", default=data_type, choices=["sigmoid", "tanh", "relu", "linear"])
    p.add_argument("--data-is-image", action="store_true")
    p.add_argument("--data-is-double", action="store_true")
    p.add_argument("--directed", action="store_true")
    p.add_argument("--data-directed", action="store_true")
    p.add_argument("--model-type", default=model_type, choices=["conv", "linear"])
    p.add_argument("--dist-type", default=dist_type, choices=["pcd", "monomer", "siamese"])
    p.add_argument("--act-type", default=act_type, choices=[None, "linear", "tanh", "sigmoid", "relu"])
    p.add_argument("--use-threshold", action="store_true", default=use_threshold)
    p.add_argument("--lr", type=float, default=lr)
    p.add_argument("--beta1", type=float, default=beta1)
    p.add_argument("--beta2", type=float, default=beta2)
    p.add_argument("--num-components", type=int, default=num_components)
    p.add_argument("--latent-size", type=int, default=latent_size)
    p.add_argument("--caffe-margin", type=float, default=caffe_margin)
    p.add_argument("--pos-weight", type=float, default=pos_weight)
    p.add_argument("--lambda-m", type=float, default=lambda_m)
    p.add_argument("--reg-const", type=float, default=reg_const)
    return p


def monomer_parser(data_name="monomer/Baby-also_viewed", data_root="parsed_data", checkpoint_root="checkpoints",
                   log_root="logs", run_tag=None, seed=633, input_shape=(4096,), batch_size=100,
                   normalize_value=1.0, num_components=2, latent_size=20, lr=0.001, beta1=0.9, beta2=0.999,
                   reg_const=0.0):
    """cfl/utils.py:510-553."""
    p = argparse.ArgumentParser()
    p.add_argument("--data-name", default=data_name)
    p.add_argument("--data-root", default=data_root)
    p.add_argument("--checkpoint-root", default=checkpoint_root)
    p.add_argument("--log-root", default=log_root)
    p.add_argument("--run-tag", default=run_tag)
    p.add_argument("--seed", type=int, default=seed)
    p.add_argument("--normalize-value", type=float, default=normalize_value)
    p.add_argument("--input-shape", nargs="+", type=int, help="shape of input", default=input_shape)
    p.add_argument("--batch-size", type=int, default=batch_size)
    p.add_argument("--num-components", type=int, default=num_components)
    p.add_argument("--latent-size", type=int, default=latent_size)
    p.add_argument("--lr", type=float, default=lr)
    p.add_argument("--beta1", type=float, default=beta1)
    p.add_argument("--beta2", type=float, default=beta2)
    p.add_argument("--reg-const", type=float, default=reg_const)
    return p
