"""python -m cfl.bin.train_dist -- cfl/bin/train_dist.py: the Monomer-style ``Dist`` model
(plain FC encoder, per-epoch checkpoint, best val accuracy -> best_acc_model)."""
import logging
import os
import shutil

from .. import variables as vs
from ..input_data import load_data_sets
from ..models.dist import construct_model
from ..utils import (IncrementalAverage, Saver, Session, dist_eval, load_best_stats, load_model, monomer_parser,
                     reduce_product)
from ._common import setup_logging

logger = logging.getLogger(__name__)


def train_loop(sess, model, data, batch_size, start_epoch, epochs, checkpoint_dir, saver):
    """cfl/bin/train_dist.py:37-117."""
    best_saver = Saver()
    nb_train = max(data.train.num_examples_labeled_pos, data.train.num_examples_labeled_neg)
    logger.warning("%d examples", nb_train)
    logger.warning("model: %s", model.get_name())
    nb_batch = nb_train // batch_size
    best_dir = os.path.join(checkpoint_dir, "best_acc_model")
    os.makedirs(best_dir, exist_ok=True)
    best_accuracy_path = os.path.join(best_dir, "best_accuracy")
    stats = load_best_stats(best_accuracy_path)
    for e in range(start_epoch, epochs):
        train_avg, val_avg = IncrementalAverage(), IncrementalAverage()
        for _ in range(nb_batch):
            out = model.train_step(*data.train.next_batch(batch_size), val_batches=data.val.next_batch(batch_size))
            train_avg.add(out["s_accuracy"])
            val_avg.add(out["val_s_accuracy"])
        saver.save(sess, os.path.join(checkpoint_dir, "model"), global_step=e)
        val_stats = dist_eval(sess, model, batch_size, data.val)
        if val_stats.accuracy > stats.best_accuracy:
            test_stats = dist_eval(sess, model, batch_size, data.test)
            logger.warning("epoch %d: current error = train: %f val: %f test: %f / auc = val: %f test: %f", e,
                           1. - train_avg.average, 1. - val_stats.accuracy, 1. - test_stats.accuracy, val_stats.auc,
                           test_stats.auc)
            stats.best_accuracy, stats.best_auc, stats.best_epoch = val_stats.accuracy, val_stats.auc, e
            best_saver.save(sess, os.path.join(best_dir, "model"), global_step=stats.best_epoch)
            with open(best_accuracy_path, "w") as outfile:
                outfile.write("{}\t{}\t{}".format(stats.best_epoch, stats.best_accuracy, stats.best_auc))
        else:
            logger.warning("epoch %d: avg error = train: %f val: %f", e, 1. - train_avg.average, 1. - val_avg.average)
    return stats


def build(args):
    vs.set_seed(args.seed)
    input_shape = tuple(args.input_shape)
    data = load_data_sets(os.path.join(args.data_root, args.data_name), reduce_product(input_shape), seed=args.seed)
    model, _ = construct_model(input_shape=input_shape, latent_size=args.latent_size,
                               normalize_value=args.normalize_value, lr=args.lr, beta1=args.beta1, beta2=args.beta2,
                               num_components=args.num_components, batch_size=args.batch_size, run_tag=args.run_tag,
                               reg_const=args.reg_const)
    return data, model


def train_monomer(args):
    data, model = build(args)
    checkpoint_dir = os.path.join(args.checkpoint_root, args.data_name, model.get_name())
    log_dir = os.path.join(args.log_root, args.data_name, model.get_name())
    for path in (checkpoint_dir, log_dir):
        if args.reset and os.path.exists(path):
            shutil.rmtree(path)
        os.makedirs(path, exist_ok=True)
    setup_logging(log_dir)
    with Session(model) as sess:
        saver, start_epoch = load_model(sess, checkpoint_dir)
        train_loop(sess=sess, model=model, data=data, batch_size=args.batch_size, start_epoch=start_epoch,
                   epochs=args.epochs, checkpoint_dir=checkpoint_dir, saver=saver)
    return model


def parse_args(argv=None):
    parser = monomer_parser()
    parser.add_argument("--epochs", type=int, default=120)
    parser.add_argument("--reset", action="store_true")
    return parser.parse_args(argv)


def main(argv=None):
    return train_monomer(parse_args(argv))


if __name__ == "__main__":
    main()
