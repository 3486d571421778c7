"""python -m cfl.bin.predict_dist -- cfl/bin/predict_dist.py: ``Dist`` scores under best_acc_model."""
import logging
import os

from ..utils import Session, monomer_parser
from ._common import setup_logging
from .predict import predict_all
from .train_dist import build

logger = logging.getLogger(__name__)


def parse_args(argv=None):
    parser = monomer_parser(batch_size=500)
    parser.add_argument("--predict-root", default="predicts")
    return parser.parse_args(argv)


def main(argv=None):
    setup_logging()
    args = parse_args(argv)
    data, model = build(args)
    checkpoint_dir = os.path.join(args.checkpoint_root, args.data_name, model.get_name())
    predict_dir = os.path.join(args.predict_root, args.data_name, model.get_name())
    with Session(model) as sess:
        predict_all(sess, model, data, args.batch_size, checkpoint_dir, predict_dir,
                    outputs=(("best_acc_model", ("predict_train_acc.txt", "predict_val_acc.txt", "predict_acc.txt")),))
    return predict_dir


if __name__ == "__main__":
    main()
