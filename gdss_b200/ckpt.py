"""Checkpoint reading for the sampler drivers (utils/loader.py:168-196, sampler.py:37-52).

The reference saves ``{model_config, params_x, params_adj, x_state_dict, adj_state_dict[, ema_x, ema_adj]}``
(trainer.py:119-127).  ``load_ckpt`` reads that file -- or the plain-dict copy of it under tests/golden/ckpt -- and
``ema_weights`` overwrites the i-th state_dict tensor with the i-th EMA shadow parameter, which is what
``ExponentialMovingAverage.copy_to`` does on the module's ``parameters()`` (utils/ema.py:37-47; every state_dict entry of
the score networks is a parameter, in the same order)."""
import os

import torch

GOLDEN_CKPT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ckpt")


def load_ckpt(path_or_name, map_location="cpu"):
    """A checkpoint path, or the name of one of the seven shipped checkpoints re-saved under tests/golden/ckpt."""
    path = path_or_name
    if not os.path.isfile(path):
        path = os.path.join(GOLDEN_CKPT_DIR, f"{path_or_name}.pt")
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        return torch.load(path, map_location=map_location, weights_only=False)      # the reference's EasyDict pickles


def strip_module(sd):
    keys = list(sd.keys())
    return {k[7:]: v for k, v in sd.items()} if keys and "module." in keys[0] else dict(sd)


def ema_weights(sd, ema):
    """state_dict with the EMA shadow parameters in place of the raw weights (sampler.py:47-52)."""
    shadow = ema["shadow_params"]
    if len(shadow) != len(sd):
        raise ValueError("EMA shadow parameters do not match the state_dict")
    for k, s in zip(sd.keys(), shadow):
        if tuple(s.shape) != tuple(sd[k].shape):
            raise ValueError(f"EMA shadow parameter of {k} has shape {tuple(s.shape)}, the state_dict has {tuple(sd[k].shape)}")
    return {k: s.detach().to(torch.float32) for k, s in zip(sd.keys(), shadow)}


def sampling_weights(ck, use_ema=None):
    """-> (x_state_dict, params_x, adj_state_dict, params_adj); EMA applied when the checkpoint's own config asks for it
    (``model_config.sample.use_ema``) or when ``use_ema`` is True."""
    sdx, sda = strip_module(ck["x_state_dict"]), strip_module(ck["adj_state_dict"])
    if use_ema is None:
        cfg = ck.get("model_config", {})
        sample = cfg.get("sample", {}) if isinstance(cfg, dict) else getattr(cfg, "sample", {})
        use_ema = bool(sample.get("use_ema", False)) if isinstance(sample, dict) else bool(getattr(sample, "use_ema", False))
        use_ema = use_ema and "ema_x" in ck
    if use_ema:
        sdx, sda = ema_weights(sdx, ck["ema_x"]), ema_weights(sda, ck["ema_adj"])
    return sdx, ck["params_x"], sda, ck["params_adj"]
