"""MVQA — mirror of vilmedic/models/mvqa/MVQA.py:14-59: VisualEncoder.forward -> adapter (Linear + LayerNorm, :23-26) ->
BertEncoder (:28-30,43) -> BertPooler (:30,47) -> Classifier (:32,48) -> loss (:34,52).
The reference asks the encoder for `output_attentions=True` and then only keeps them for post-processing; attention
probabilities are not materialised here (fused attention), `attentions` is not part of the returned dict either way."""
import torch
import torch.nn as nn

from ...blocks.classifier import *  # noqa: F401,F403
from ...blocks.classifier.evaluation import evaluation
from ...blocks.losses import *  # noqa: F401,F403
from ...blocks.vision import *  # noqa: F401,F403
from ...cfgutil import to_attrdict
from ...nn import BertEncoderB200, TanhFn, bert_config, native_layernorm, native_linear, set_arena_root


class _Pooler(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.dense = nn.Linear(d, d)


class MVQA(nn.Module):
    def __init__(self, cnn, classifier, adapter, transformer, loss, **kwargs):
        super().__init__()
        cnn, classifier, adapter = to_attrdict(cnn), to_attrdict(classifier), to_attrdict(adapter)
        transformer, loss = to_attrdict(transformer), to_attrdict(loss)
        cnn_func, loss_func, classifier_func = cnn.pop("proto"), loss.pop("proto"), classifier.pop("proto")
        self.cnn = eval(cnn_func)(**cnn)
        self.adapter = nn.Sequential(nn.Linear(adapter.pop("input_size"), adapter.pop("output_size")),
                                     nn.LayerNorm(transformer.hidden_size, eps=transformer.layer_norm_eps))
        conf = bert_config(**dict(transformer))
        self.transformer = BertEncoderB200(conf)
        self.pooler = _Pooler(conf.hidden_size)
        self.classifier = eval(classifier_func)(**classifier)
        self.loss_func = eval(loss_func)(**loss)
        self.eval_func = evaluation
        set_arena_root(self)

    def forward(self, images, labels=None, from_training=True, iteration=None, epoch=None, **kwargs):
        out = self.cnn(images.cuda(non_blocking=True))                       # [B, S, C] bf16
        B, S, C = out.shape
        h = native_linear(self.adapter[0], out.reshape(B * S, C), self)
        h = native_layernorm(self.adapter[1], h, self)
        h = self.transformer(h.view(B, S, -1))
        first = h[:, 0].contiguous()
        pooled = TanhFn.apply(native_linear(self.pooler.dense, first, self, out_dtype=torch.float32))
        out = self.classifier(pooled)
        loss = torch.tensor(0.)
        if from_training:
            loss = self.loss_func(out, labels.cuda(non_blocking=True), **kwargs)
        return {"loss": loss, "output": out, "answer": torch.argmax(out, dim=-1)}
