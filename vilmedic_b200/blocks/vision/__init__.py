from .visual_encoder import *  # noqa: F401,F403  (mirrors vilmedic/blocks/vision/__init__.py)
