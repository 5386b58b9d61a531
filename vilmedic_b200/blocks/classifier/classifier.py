"""Classifier — mirror of vilmedic/blocks/classifier/classifier.py:4-15 (Dropout -> Linear), the Linear on the tcgen05 GEMM."""
import torch
import torch.nn as nn

from ...nn import DropoutFn, native_linear


class Classifier(nn.Module):
    def __init__(self, input_size, num_classes, dropout=0., **kwargs):
        super().__init__()
        self.classifier = nn.Sequential(nn.Linear(in_features=input_size, out_features=num_classes))
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, input):
        x = input
        if self.dropout.p > 0 and self.training and torch.is_grad_enabled():
            x = DropoutFn.apply(x.to(torch.bfloat16).contiguous(), self.dropout.p)
        return native_linear(self.classifier[0], x.reshape(-1, x.shape[-1]), self, out_dtype=torch.float32).view(*x.shape[:-1], -1)
