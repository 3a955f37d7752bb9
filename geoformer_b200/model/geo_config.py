"""Inference-time GeoFormer options — same keys and values as the reference's
``model/geo_config.py:9-19`` (``default_cfg``), kept as a plain mutable dict because the
immatch wrapper mutates it in place (eval_tool/immatch/modules/geoformer.py:23-25)."""

default_cfg = {
    "layer_names": ["self", "cross"] * 2,
    "nhead": 4,
    "coarse_thr": 0.2,
    "fine_temperature": 0.1,
    "fine_thr": 0.1,
    "window_size": 5,
    "topk": 1,
}


def get_cfg_model():
    """Reference API (geo_config.py:23-27): a private copy of the defaults (upper-case keys there; plain dict here)."""
    import copy
    return copy.deepcopy(default_cfg)
