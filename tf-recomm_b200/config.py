"""Module-level constants under the reference's names (config.py in jilljenn/TF-recomm: DIM, EPOCH_MAX,
LEARNING_RATE, LAMBDA_REG, DISCRETE, DEVICE, PREFIX, BASE_DIR, ARTICLE_FOLDER) plus the four the reference's
drivers use but never define -- BATCH_SIZE, USER_NUM, ITEM_NUM, NB_CLASSES (svd_train_val.py:24,47,80; they
live in per-dataset config.yml files, fm_fraction.py:28-32) -- and the switch between the two models the
repository documents.  Every value can be overridden with an environment variable TFR_<NAME> or, for a
dataset, with `load_dataset_config(path)`.

MODEL_VARIANT:
  "readme"  y = mu + b_u + b_i + <p_u, q_i>, squared error + LAMBDA_REG*(|p_u|^2+|q_i|^2), Adam
            (README.md:31-39, doc/graph_svd.png -- the path BASELINE.json's north_star names)
  "fork"    the code as written: |q_i| in the dot, sigmoid cross-entropy, L2 on biases too, plain SGD
            (ops.py:44,85-89,125,145)
"""
import os


def _env(name, default, cast):
    raw = os.environ.get("TFR_" + name)
    return default if raw is None else cast(raw)


MODEL_VARIANT = _env("MODEL_VARIANT", "readme", str)

# README run: dim 15, lr 1e-3, reg 0.05, batch 1000 on MovieLens-1M shapes.
DIM = _env("DIM", 15, int)
EPOCH_MAX = _env("EPOCH_MAX", 100, int)
LEARNING_RATE = _env("LEARNING_RATE", 1e-3, float)
LAMBDA_REG = _env("LAMBDA_REG", 0.05, float)
BATCH_SIZE = _env("BATCH_SIZE", 1000, int)
USER_NUM = _env("USER_NUM", 6040, int)
ITEM_NUM = _env("ITEM_NUM", 3952, int)
NB_CLASSES = _env("NB_CLASSES", 2, int)
# DISCRETE selects the driver's metric branch (svd_train_val.py:79,99): binary outcomes -> ACC/AUC/NLL,
# ratings -> RMSE.  The README model is the RMSE branch.
DISCRETE = _env("DISCRETE", MODEL_VARIANT == "fork", lambda s: s.lower() in ("1", "true", "yes"))

DEVICE = _env("DEVICE", "/gpu:0", str)   # kept for signature compatibility; the tables always live in HBM
PREFIX = ""
BASE_DIR = _env("BASE_DIR", os.getcwd(), str)
ARTICLE_FOLDER = _env("ARTICLE_FOLDER", os.path.join(BASE_DIR, "article"), str)
SEED = 13575  # svd_train_val.py:15


def load_dataset_config(config_file):
    """Apply a dataset's config.yml (keys USER_NUM, ITEM_NUM, NB_CLASSES, BATCH_SIZE) to this module."""
    from . import dataio
    cfg = dataio.get_config(config_file) or {}
    g = globals()
    for k in ("USER_NUM", "ITEM_NUM", "NB_CLASSES", "BATCH_SIZE"):
        if k in cfg:
            g[k] = int(cfg[k])
    return cfg
