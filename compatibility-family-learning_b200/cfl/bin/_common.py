"""Shared set-up of the cfl.bin entry points: data sets, normalisers and the model from parsed flags."""
import logging
import os

from ..input_data import load_data_sets
from ..ops import dist_normalizer
from ..utils import reduce_product

LOG_FORMAT = "%(asctime)s [%(levelname)-5.5s] [%(name)s]  %(message)s"


def setup_logging(log_dir=None):
    """Warnings to ``<log_dir>/log.log`` and the console (cfl/bin/train.py:236-247)."""
    root = logging.getLogger()
    logging.basicConfig(format=LOG_FORMAT, level=logging.WARNING)
    if log_dir:
        path = os.path.abspath(os.path.join(log_dir, "log.log"))
        if not any(getattr(h, "baseFilename", None) == path for h in root.handlers):
            fh = logging.FileHandler(path)
            fh.setLevel(logging.WARNING)
            fh.setFormatter(logging.Formatter(LOG_FORMAT))
            root.addHandler(fh)


def build_cfl(args, data_switch=False):
    """What cfl/bin/train.py:104-219 and cfl/bin/predict.py:34-120 do before opening the session."""
    from ..models.cfl import construct_model
    from .. import variables as vs
    vs.set_seed(args.seed)
    if args.data_is_image:
        raise NotImplementedError("image records feed the conv / generation half (out of scope)")
    feature_shape = args.latent_shape if args.data_is_double else args.input_shape
    data = load_data_sets(os.path.join(args.data_root, args.data_name), reduce_product(feature_shape),
                          directed=args.directed or args.data_directed, data_switch=data_switch, seed=args.seed)
    (data_normalizer, _, _, _, latent_normalizer) = dist_normalizer(
        input_shape=args.input_shape, ae_shape=None, data_scale=args.data_scale, data_mean=args.data_mean,
        data_norm=args.data_norm, latent_norm=args.latent_norm, data_type=args.data_type)
    model, aux = construct_model(
        input_shape=args.input_shape, latent_shape=args.latent_shape, batch_size=args.batch_size,
        latent_size=args.latent_size, num_components=args.num_components, model_type=args.model_type,
        dist_type=args.dist_type, act_type=args.act_type, data_type=args.data_type,
        use_threshold=args.use_threshold, pos_weight=args.pos_weight, caffe_margin=args.caffe_margin,
        lambda_m=args.lambda_m, reg_const=args.reg_const, directed=args.directed, lr=args.lr, beta1=args.beta1,
        beta2=args.beta2, data_normalizer=data_normalizer, latent_normalizer=latent_normalizer,
        data_norm=args.data_norm, latent_norm=args.latent_norm, run_tag=args.run_tag)
    return data, model, aux
