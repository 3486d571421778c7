"""python -m cfl.bin.train -- cfl/bin/train.py for the distance model (no GAN stage)."""
import logging
import os
import shutil

from ..utils import Saver, Session, dist_check_args, dist_parser, load_model
from ._common import build_cfl, setup_logging

logger = logging.getLogger(__name__)


def train_dist(args):
    data, model, _ = build_cfl(args, data_switch=args.data_switch)
    checkpoint_dir = os.path.join(args.checkpoint_root, args.data_name, model.get_name())
    log_dir = os.path.join(args.log_root, args.data_name, model.get_name())
    for path in (checkpoint_dir, log_dir):
        if args.reset and os.path.exists(path):
            shutil.rmtree(path)
        os.makedirs(path, exist_ok=True)
    setup_logging(log_dir)
    logger.warning("run with %s", model.get_name())
    best_dir = os.path.join(checkpoint_dir, "best_model")
    best_acc_dir = os.path.join(checkpoint_dir, "best_acc_model")
    os.makedirs(best_dir, exist_ok=True)
    os.makedirs(best_acc_dir, exist_ok=True)
    with Session(model) as sess:
        saver, start_iter = load_model(sess, checkpoint_dir)
        model.train(sess=sess, data=data, start_iter=start_iter, epochs=args.epochs, post_epochs=args.post_epochs,
                    best_dir=best_dir, best_acc_dir=best_acc_dir, checkpoint_dir=checkpoint_dir,
                    eval_epochs=args.eval_epochs, disable_eval=args.disable_eval, saver=saver, best_saver=Saver(),
                    best_acc_saver=Saver(), save_iters=args.save_iters, cuda_graph=args.cuda_graph)
    return model


def parse_args(argv=None):
    parser = dist_parser(batch_size=100)
    parser.add_argument("--load-pre-weights", action="store_true")
    parser.add_argument("--epochs", type=int, default=120)
    parser.add_argument("--save-iters", type=int)
    parser.add_argument("--data-switch", action="store_true")
    parser.add_argument("--post-epochs", type=int, default=100)
    parser.add_argument("--eval-epochs", type=int, default=1)
    parser.add_argument("--disable-eval", action="store_true")
    parser.add_argument("--reset", action="store_true")
    parser.add_argument("--cuda-graph", action="store_true",
                        help="replay each train step as one CUDA graph (linear models, single process)")
    args = parser.parse_args(argv)
    dist_check_args(args)
    return args


def main(argv=None):
    return train_dist(parse_args(argv))


if __name__ == "__main__":
    main()
