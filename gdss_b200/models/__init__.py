from .score_networks import ScoreNetworkA, ScoreNetworkX, ScoreNetworkX_GMH, load_model, load_model_from_ckpt  # noqa: F401
