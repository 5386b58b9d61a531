"""RRG — mirror of vilmedic/models/rrg/RRG.py:10-52 (image encoder -> cross-attending report decoder)."""
import torch
import torch.nn as nn

from ...blocks.huggingface.decoder.decoder_model import DecoderModel
from ...blocks.huggingface.decoder.evaluation import evaluation
from ...blocks.vision import *  # noqa: F401,F403  (the `proto` string is eval()-ed against these names, RRG.py:20)
from ...cfgutil import to_attrdict
from ...nn import set_arena_root


class RRG(nn.Module):
    def __init__(self, decoder, cnn, dl=None, **kwargs):
        super().__init__()
        decoder = to_attrdict(decoder)
        cnn = to_attrdict(cnn)
        if dl:
            decoder.vocab_size = dl.dataset.seq.tokenizer.vocab_size
        self.dec = DecoderModel(decoder)
        self.enc = eval(cnn.pop("proto"))(**cnn)
        self.eval_func = evaluation
        set_arena_root(self)

    def forward(self, input_ids, attention_mask, images, images_mask=None, encoder_outputs=None,
                encoder_attention_mask=None, epoch=None, iteration=None, **kwargs):
        input_ids = input_ids.cuda(non_blocking=True)
        attention_mask = attention_mask.cuda(non_blocking=True)
        if encoder_outputs is None:
            encoder_outputs, encoder_attention_mask = self.encode(images, images_mask, **kwargs)
        return self.dec(input_ids=input_ids, attention_mask=attention_mask, encoder_outputs=encoder_outputs,
                        encoder_attention_mask=encoder_attention_mask, **kwargs)

    def encode(self, images, images_mask=None, **kwargs):
        return self.enc.encode(images, images_mask, **kwargs)

    def __repr__(self):
        n = sum(p.numel() for p in self.parameters())
        return "model: RRG\n(enc):%s\n(dec):%s\n{'n_params': %d}\n" % (self.enc, self.dec, n)
