def get_cls_params(cls):
    return (getattr(cls, "_req_params", {}), getattr(cls, "_opt_params", {}), getattr(cls, "_single_params", []),
            getattr(cls, "_takes_rng", False))
