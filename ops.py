"""Drop-in for the reference's ops.py: same module name and entry points, backed by tf-recomm_b200."""
from tf_recomm_b200.ops import *  # noqa: F401,F403
from tf_recomm_b200.ops import inference_svd, optimization, sigmoid  # noqa: F401
