"""Config access helpers: the reference passes OmegaConf DictConfig nodes (attribute + item access, .pop);
plain dicts and simple namespaces are accepted as well (omegaconf is not a dependency of this package)."""


class AttrDict(dict):
    """dict with attribute access and `pop`, the subset of DictConfig the reference's blocks rely on
    (vilmedic/models/rrg/RRG.py:15-20, vilmedic/blocks/huggingface/decoder/decoder_model.py:14-26)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def to_attrdict(cfg):
    if cfg is None:
        return AttrDict()
    if isinstance(cfg, AttrDict):
        return cfg
    if isinstance(cfg, dict):
        return AttrDict({k: (to_attrdict(v) if isinstance(v, dict) else v) for k, v in cfg.items()})
    if hasattr(cfg, "items"):  # DictConfig
        out = AttrDict()
        for k, v in cfg.items():
            out[k] = to_attrdict(v) if hasattr(v, "items") else v
        return out
    if hasattr(cfg, "__dict__"):
        return to_attrdict(dict(vars(cfg)))
    raise TypeError("unsupported config node: %r" % type(cfg))


def cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict) or hasattr(cfg, "keys"):
        try:
            return cfg[key] if key in cfg else default
        except Exception:
            return default
    return getattr(cfg, key, default)


def cfg_pop(cfg, key, default=None):
    if hasattr(cfg, "pop"):
        try:
            return cfg.pop(key)
        except KeyError:
            return default
    v = getattr(cfg, key, default)
    if hasattr(cfg, key):
        delattr(cfg, key)
    return v
