"""Thin orchestration helpers mirroring the reference entry points that touch the hot path (vilmedic/executors/__init__.py:1-6):
`create_model` (how a config's `model.proto` string becomes a model), config loading with `includes`, the optimizer factory and
a synthetic data loader of the reference batch shape.  Trainor / Validator themselves stay with the reference (north_star)."""
from .utils import (SyntheticLoader, create_data_loader, create_model, create_optimizer, load_config,  # noqa: F401
                    vilmedic_state_dict_versioning)
