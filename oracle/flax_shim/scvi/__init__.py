class _RK:
    X_KEY = "X"
    BATCH_KEY = "batch"
    LABELS_KEY = "labels"
    CONT_COVS_KEY = "extra_continuous_covs"
    INDICES_KEY = "ind_x"
    SIZE_FACTOR_KEY = "size_factor"


REGISTRY_KEYS = _RK()
