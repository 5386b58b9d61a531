from vilmedic_b200.blocks.vision import *  # noqa: F401,F403
from vilmedic_b200.blocks.vision import VisualEncoder, get_network  # noqa: F401
