"""Reference checkpoint interchange.

The reference saves `{"model": state_dict, "optimizer": ..., "training_scheduler": ..., "config": ..., "scaler": ...,
"__version__": ...}` (vilmedic/executors/trainor.py:193-203, utils.py:250-263) and, when loading, strips the
`module.` prefix that `nn.DataParallel` adds and renames the pre-1.3.2 visual-encoder keys
(`vilmedic_state_dict_versioning`, vilmedic/executors/utils.py:26-34) before `load_state_dict(strict=True)` (utils.py:113-119).
The kernel towers keep the HF parameter names and shapes, so a reference checkpoint loads into them unchanged once the same
key normalisation has been applied — and `reference_state_dict(model)` writes one the reference can load back.
Host-side dictionary work only; no arithmetic.
"""
import torch


def normalize_reference_keys(params, version=None):
    """Same renames, in the same order, as vilmedic_state_dict_versioning (vilmedic/executors/utils.py:26-34)."""
    params = {k.replace("module.", ""): v for k, v in params.items()}
    if version is None or version < "1.3.2":
        params = {k.replace("enc.0.cnn.", "enc.model."): v for k, v in params.items()}
        params = {k.replace("enc.1.weight", "enc.visual_projection.weight"): v for k, v in params.items()}
        params = {k.replace("enc.1.bias", "enc.visual_projection.bias"): v for k, v in params.items()}
    return params


def load_reference_checkpoint(model, ckpt, strict=True):
    """ckpt: path to a reference `.pth` or the already loaded dict.  Raises like the reference when "model" is missing."""
    if isinstance(ckpt, (str, bytes)):
        # Reference checkpoints carry non-tensor entries ('config': an OmegaConf DictConfig, the scheduler object, ... —
        # vilmedic/executors/trainor.py:194-199), which torch >= 2.6's weights_only=True default refuses to unpickle.  They are
        # the user's own training artefacts, i.e. trusted input, exactly as the reference treats them (torch.load, utils.py:241).
        # If a pickled class is not importable here (omegaconf absent), fall back to reading only the tensors.
        try:
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
        except (ModuleNotFoundError, AttributeError, ImportError):
            ckpt = _load_tensors_only(ckpt)
    if "model" not in ckpt:
        raise KeyError('This checkpoint is not valid. Key "model" is missing from dict.')
    params = normalize_reference_keys(ckpt["model"], ckpt.get("__version__", None))
    return model.load_state_dict(params, strict=strict)


class _Opaque:
    """Placeholder for pickled objects whose class cannot be imported here (e.g. omegaconf.DictConfig)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


def _load_tensors_only(path):
    import pickle

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ModuleNotFoundError, AttributeError, ImportError):
                return _Opaque

    class _Module:
        Unpickler = _Unpickler
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
        __name__ = "pickle"

    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Module)


def reference_state_dict(model, version="1.3.3", **extra):
    """A checkpoint dict in the reference's layout (CPU fp32 tensors) holding this model's parameters."""
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    out = {"model": sd, "__version__": version}
    out.update(extra)
    return out
