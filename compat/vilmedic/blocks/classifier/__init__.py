from vilmedic_b200.blocks.classifier import Classifier  # noqa: F401
