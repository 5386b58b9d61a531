from vilmedic_b200.blocks.huggingface.encoder.encoder_model import EncoderModel  # noqa: F401
