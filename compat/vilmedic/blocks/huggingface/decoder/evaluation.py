from vilmedic_b200.blocks.huggingface.decoder.evaluation import *  # noqa: F401,F403
from vilmedic_b200.blocks.huggingface.decoder.evaluation import evaluation  # noqa: F401
