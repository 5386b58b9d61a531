"""Mirror of vilmedic/blocks/losses/__init__.py: the hot-path losses on the B200 kernels + every torch.nn loss by name
(the reference star-exports torch.nn.modules.loss so that `proto: BCEWithLogitsLoss` resolves, losses/__init__.py:6)."""
from torch.nn.modules.loss import *  # noqa: F401,F403

from .contrastive import ConVIRTLoss, GLoRIAGlobalLoss, InfoNCELoss  # noqa: F401
from .gloria import GLoRIALoss, cosine_similarity, gloria_attention_fn, global_loss, local_loss  # noqa: F401
from .label_smoothing import LabelSmoothingCrossEntropy  # noqa: F401
