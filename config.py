"""Drop-in for the reference's config.py (`from config import *` in its drivers)."""
from tf_recomm_b200.config import *  # noqa: F401,F403
from tf_recomm_b200.config import (ARTICLE_FOLDER, BASE_DIR, BATCH_SIZE, DEVICE, DIM, DISCRETE, EPOCH_MAX,  # noqa: F401
                                   ITEM_NUM, LAMBDA_REG, LEARNING_RATE, MODEL_VARIANT, NB_CLASSES, PREFIX, SEED,
                                   USER_NUM)
