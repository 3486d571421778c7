"""python -m cfl.bin.predict -- cfl/bin/predict.py: scores of every labelled pair of train / val /
test under the best-AUC and the best-accuracy checkpoints, in the files evaluate_total reads."""
import logging
import os

from ..utils import Session, dist_check_args, dist_parser, dist_predict, load_model
from ._common import build_cfl, setup_logging

logger = logging.getLogger(__name__)

OUTPUTS = (("best_model", ("predict_train.txt", "predict_val.txt", "predict.txt")),
           ("best_acc_model", ("predict_train_acc.txt", "predict_val_acc.txt", "predict_acc.txt")))


def predict_all(sess, model, data, batch_size, checkpoint_dir, predict_dir, outputs=OUTPUTS):
    for sub, names in outputs:
        _, start = load_model(sess, os.path.join(checkpoint_dir, sub))
        if start == 0:
            raise FileNotFoundError("no checkpoint under %s" % os.path.join(checkpoint_dir, sub))
        for split, name in zip((data.train, data.val, data.test), names):
            dist_predict(sess=sess, model=model, data=split, batch_size=batch_size, predict_dir=predict_dir,
                         output_name=name)


def parse_args(argv=None):
    parser = dist_parser(batch_size=500)
    parser.add_argument("--predict-root", default="predicts")
    args = parser.parse_args(argv)
    dist_check_args(args)
    return args


def main(argv=None):
    setup_logging()
    args = parse_args(argv)
    data, model, _ = build_cfl(args)
    checkpoint_dir = os.path.join(args.checkpoint_root, args.data_name, model.get_name())
    predict_dir = os.path.join(args.predict_root, args.data_name, model.get_name())
    logger.warning("run with %s", model.get_name())
    with Session(model) as sess:
        predict_all(sess, model, data, args.batch_size, checkpoint_dir, predict_dir)
    return predict_dir


if __name__ == "__main__":
    main()
