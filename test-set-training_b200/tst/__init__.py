"""tst -- B200-native TsT Random-Forest diagnostic (host side).  See DESIGN.md."""
from ._lib import FOLD_INACTIVE, FOLD_TRAIN_ONLY, TstError  # noqa: F401
from .api import (Engine, IBPResult, TsTForest, TsTResult, get_engine, ibp, ibp_vmin, make_params,  # noqa: F401
                  resolve_max_features, run_tst, shard_trees)
from .features import extract_features, gini_importances, to_sklearn_arrays  # noqa: F401,E402
