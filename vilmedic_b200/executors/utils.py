"""create_model / create_optimizer / config loading — mirror of vilmedic/executors/utils.py:26-34,65-137 and of the `includes`
merge in bin/utils.py:93-139, without OmegaConf (not installed here; plain dict nodes with attribute access, cfgutil.AttrDict).

`create_model` resolves `config.model.proto` by `eval()` against `vilmedic_b200.models` exactly as the reference evaluates it
against `vilmedic.models` (utils.py:17,105-110), passes the remaining keys as kwargs together with `dl`, `logger`,
`from_training`, applies the checkpoint key versioning and `load_state_dict(strict=True)` (utils.py:113-119), and moves the
model to the GPU when one is present (the reference calls `.cuda()` unconditionally; here a CPU-only host — the test box —
keeps the constructed model on the CPU so that configs can be validated without a GPU; nothing computes there).
One process drives one GPU (torch.distributed): the reference's nn.DataParallel wrapping (utils.py:128-133) is
intentionally not reproduced.
"""
import copy
import os
import re

import torch
import yaml

from .. import synth
from ..cfgutil import AttrDict, to_attrdict
from ..checkpoint import normalize_reference_keys as vilmedic_state_dict_versioning  # noqa: F401
from ..models import *  # noqa: F401,F403  (names `model.proto` is eval()-ed against)
from .. import optim as _optim

_NUM = re.compile(r"^-?(\d+\.?\d*|\d*\.?\d+)([eE][+-]?\d+)?$")


def _convert_numeric_strings(obj):
    """bin/utils.py:35-66: YAML reads `1e-05` / `5e-5` as strings; the reference converts them after merging."""
    if isinstance(obj, str):
        if _NUM.match(obj.strip()):
            try:
                return int(obj) if ("." not in obj and "e" not in obj.lower()) else float(obj)
            except ValueError:
                return obj
        return obj
    if isinstance(obj, dict):
        return {k: _convert_numeric_strings(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_convert_numeric_strings(v) for v in obj]
    return obj


def _merge(base, over):
    """OmegaConf.merge semantics for the node types the configs use: dicts merge recursively, everything else is replaced."""
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def _set_dotted(d, key, value):
    parts = key.split(".")
    for p in parts[:-1]:
        if not isinstance(d.get(p), dict):
            d[p] = {}
        d = d[p]
    d[parts[-1]] = value


def load_config(path, overrides=()):
    """Config file + `includes` + `key.sub=value` overrides -> attribute-access dict (bin/utils.py:93-139)."""
    with open(path) as f:
        config = yaml.safe_load(f) or {}
    includes = config.get("includes", [])
    if not isinstance(includes, (list, tuple)):
        raise AttributeError("Includes must be a list, {} provided".format(type(includes)))
    merged = {}
    for inc in includes:
        if not os.path.exists(inc):
            inc = os.path.join(os.path.dirname(path), inc)
        merged = _merge(merged, dict(load_config(inc)))
    config = _merge(merged, config)
    for arg in overrides:
        key, _, value = arg.partition("=")
        _set_dotted(config, key, yaml.safe_load(value) if value != "" else None)
    return to_attrdict(_convert_numeric_strings(_plain(config)))


def _plain(node):
    if isinstance(node, dict):
        return {k: _plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [_plain(v) for v in node]
    return node


def get(config, mode):
    """Per-executor view of the config (bin/utils.py:142-150): the executor's own block + every non-executor top-level key."""
    exec_config = copy.deepcopy(config[mode])
    for att in list(config.keys()):
        if att not in ("trainor", "validator", "ensemblor"):
            exec_config[att] = config[att]
    return to_attrdict(exec_config)


class _Logger:
    def settings(self, msg):
        pass

    info = warning = critical = error = settings


def create_model(config, dl, logger=None, from_training=True, state_dict=None, from_accelerate=False):
    logger = logger or _Logger()
    config_copy = copy.deepcopy(config.model)
    if "proto" not in config_copy:
        raise ValueError("config.model.proto is required")
    proto = config_copy.get("proto")
    config_dict = {k: v for k, v in config_copy.items() if k != "proto"}
    model = eval(proto)(**config_dict, dl=dl, logger=logger, from_training=from_training)
    logger.settings("Model {} created".format(type(model).__name__))
    if state_dict is not None:
        if "model" not in state_dict:
            raise KeyError('This checkpoint is not valid. Key "model" is missing from dict.')
        params = vilmedic_state_dict_versioning(state_dict["model"], state_dict.get("__version__", None))
        model.load_state_dict(params, strict=True)
        logger.info("Model state loaded")
    else:
        logger.info(model)
    if not from_accelerate and torch.cuda.is_available():
        model = model.cuda()
    return model


def create_optimizer(config, logger, model, state_dict=None):
    """vilmedic/executors/utils.py:65-95 on the fused kernels: `config.optimizer` names a torch.optim class (RAdam / Adam / AdamW
    have kernels, anything else raises NotImplementedError like an unknown name does in the reference); `model` replaces the
    reference's `model.parameters()` argument because the fused step owns the flat arena."""
    if "optim_params" not in config or "lr" not in config.optim_params:
        raise ValueError("config.optim_params.lr is required")
    if "optimizer" not in config:
        raise ValueError("config.optimizer is required")
    return _optim.create_optimizer(config.optimizer, model, state_dict=state_dict, **dict(config.optim_params))


# ------------------------------------------------------------------------------------------------ synthetic data
class _Tokenizer:
    """The attributes the models / evaluation read from the dataset tokenizer (vilmedic/datasets/base/TextDataset.py:86-119,
    vocabulary order of datasets/base/utils.py:24-25: [CLS] [PAD] [SEP] [UNK] [MASK], then words)."""

    def __init__(self, vocab_size):
        self.vocab_size = vocab_size
        self.cls_token, self.pad_token, self.sep_token, self.unk_token, self.mask_token = "[CLS]", "[PAD]", "[SEP]", "[UNK]", "[MASK]"
        self.cls_token_id, self.pad_token_id, self.sep_token_id, self.unk_token_id, self.mask_token_id = 0, 1, 2, 3, 4
        self.vocab = {self.cls_token: 0, self.pad_token: 1, self.sep_token: 2, self.unk_token: 3, self.mask_token: 4}

    def get_vocab(self):
        v = dict(self.vocab)
        v.update({"w%d" % i: i for i in range(5, self.vocab_size)})
        return v

    def decode(self, ids, skip_special_tokens=True, clean_up_tokenization_spaces=False):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return " ".join("w%d" % i for i in ids if not (skip_special_tokens and i < 5))


class _SynthDataset:
    def __init__(self, config, n, seed):
        ds = config.dataset
        self.proto = ds.get("proto", "ImSeq")
        self.n, self.seed = n, seed
        self.crop = ds.image.get("crop", 224)
        if self.proto == "ImSeq":
            self.tokenizer = _Tokenizer(ds.seq.vocab_size)
            self.tokenizer_max_len = ds.seq.tokenizer_max_len
            self.seq = self                  # dl.dataset.seq.tokenizer (RRG.py:15-16) and dl.dataset.tokenizer (evaluation.py:27)
        else:
            self.num_classes = ds.label.num_classes

    def __len__(self):
        return self.n


class SyntheticLoader:
    """Batches of the reference collate's shape (ImSeq: TextDataset.py:110-119 + ImageDataset.py:25-60; ImLabel: LabelDataset.py:64-69),
    seeded as SURVEY.md §8d prescribes; CPU tensors, pinned when a GPU is present."""

    def __init__(self, config, batch_size, n_batches=2, seed=1234):
        self.dataset = _SynthDataset(config, batch_size * n_batches, seed)
        self.batch_size, self.n_batches, self.seed = batch_size, n_batches, seed

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        d = self.dataset
        for i in range(self.n_batches):
            if d.proto == "ImSeq":
                b = synth.rrg_batch(self.batch_size, d.tokenizer_max_len, d.tokenizer.vocab_size, image_size=d.crop, seed=self.seed + i)
            else:
                g = torch.Generator().manual_seed(self.seed + i)
                b = {"images": torch.randn(self.batch_size, 3, d.crop, d.crop, generator=g), "images_mask": None,
                     "labels": torch.randint(0, d.num_classes, (self.batch_size,), generator=g)}
            if torch.cuda.is_available():
                b = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
            yield b


def create_data_loader(config, split, logger=None, called_by_validator=False, called_by_ensemblor=False, from_accelerate=False,
                       n_batches=2):
    """Synthetic stand-in for vilmedic/executors/utils.py:140-200 (datasets / tokenizers are out of scope): only
    configs whose dataset blocks say `synthetic: true` are served; batch size comes from the executor block like the reference."""
    if not config.dataset.image.get("synthetic", False):
        raise NotImplementedError("real datasets stay with the reference's data pipeline; this loader serves `synthetic: true` configs")
    return SyntheticLoader(config, config.batch_size, n_batches=n_batches)
