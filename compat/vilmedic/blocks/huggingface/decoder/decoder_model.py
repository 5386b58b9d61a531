from vilmedic_b200.blocks.huggingface.decoder.decoder_model import DecoderModel  # noqa: F401
