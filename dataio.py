"""Drop-in for the reference's dataio.py: same module name and entry points, backed by tf-recomm_b200."""
from tf_recomm_b200.dataio import *  # noqa: F401,F403
from tf_recomm_b200.dataio import (OneEpochIterator, ShuffleIterator, build_new_paths, build_paths,  # noqa: F401
                                   get_config, get_data, get_legend, get_new_data, prepare_folder, read_process)
