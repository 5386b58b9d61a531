from vilmedic_b200.blocks.vision.visual_encoder import *  # noqa: F401,F403
from vilmedic_b200.blocks.vision.visual_encoder import VisualEncoder, get_network  # noqa: F401
