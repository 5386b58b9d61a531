from vilmedic_b200.blocks.losses import *  # noqa: F401,F403
from vilmedic_b200.blocks.losses import (ConVIRTLoss, GLoRIALoss, InfoNCELoss, LabelSmoothingCrossEntropy, cosine_similarity,  # noqa: F401
                                         gloria_attention_fn)
