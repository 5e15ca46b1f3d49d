"""tst -- B200-native TsT Random-Forest diagnostic (host side).  See DESIGN.md."""
from ._lib import FOLD_INACTIVE, FOLD_TRAIN_ONLY, TstError  # noqa: F401
from .api import (Engine, IBPResult, TsTForest, TsTRegressionResult, TsTResult, get_engine, ibp, ibp_vmin,  # noqa: F401
                  make_params, quantise_target, quantise_target16, resolve_max_features, run_tst, run_tst_regression, shard_trees,
                  votes16_to_value, votes_to_value)
from .features import extract_features, gini_importances, path_contributions, records_from, to_sklearn_arrays  # noqa: F401,E402
from .interop import ExportedForest, load_forest  # noqa: F401,E402
