"""tf-recomm_b200: B200 (sm_100a) implementation of TF-recomm's matrix-factorization train step.

The directory name carries a hyphen (repo convention); import it as `tf_recomm_b200` through the alias module
at the repo root.  Submodules: _lib (ctypes binding of libtfrecomm.so), engine (device state + step calls),
ops / dataio / config / session (host-side mirror of the reference's Python surface), init, synthetic.
"""
import os as _os

# More hardware work queues than the default 8: the step forks onto side streams, and two streams that hash to
# the same queue serialise behind each other.  Only effective if set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__all__ = ["_lib", "engine", "init"]
