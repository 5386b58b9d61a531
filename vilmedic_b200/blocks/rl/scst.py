"""SCST (self-critical sequence training) — mirror of vilmedic/blocks/rl/SCST.py:12-190 on the decode-step kernels.

Reference flow: a greedy rollout without grad (`forward_greedy` :112-129) gives the baseline reward; a sampled rollout
(`generate(do_sample=True, top_k, bad_words_ids=[[pad],[bos]], forced_eos_token_id=True)` :139-153) gives the sequences whose
log-probabilities — log_softmax of the PROCESSED step scores, gathered at the sampled tokens (:155-158) — are weighted by
(reward_sampling - reward_greedy) in `scst_loss` (:12-44).

Here both rollouts run on the device-side search (blocks/huggingface/decoder/beam.py, one CUDA graph replay per token).  The
reference back-propagates through the incremental generate loop; the same gradient is obtained here by ONE teacher-forced pass of
the decoder over the sampled sequence with the same logit processors applied (vlm_logits_filter) — log p(y_t | y_<t) of the
filtered distribution is the same function of the parameters either way — and the reward weights folded into the fused
softmax-CE kernel as per-row gradient weights (`sequence_log_probs`).  Scorers (ROUGE, BLEU, CheXbert, ...: CPU string metrics) are
outside the hot path: pass callables `scorer(refs, hyps) -> per-sample rewards`, or names resolvable in
`vilmedic.blocks.scorers.scores.REWARD_COMPLIANT` when the reference package is importable.
"""
import json

import torch
import torch.nn as nn

from ... import ops
from ...arena import get_arena
from ...nn import _root_of


class _SeqLogProbFn(torch.autograd.Function):
    """log p(target_t) under softmax(filtered logits) for every row, with the gradient computed by the fused CE kernel.
    The logits of this pass are kept in fp32 (LM-head GEMM with fp32 output): the top-k processor compares scores with the k-th
    largest one, and bf16-rounded logits tie there far too often (HF applies it to fp32 scores as well)."""

    @staticmethod
    def forward(ctx, h, anchor, targets, head, arena, bad_ids, top_k):
        E = head.decoder.weight
        V, D = E.shape
        Vp = (V + 7) // 8 * 8
        R = h.shape[0]
        bias = arena.fp32(head.bias)
        if Vp != V:
            bp = torch.zeros(Vp, device=h.device, dtype=torch.float32)
            bp[:V] = bias
            bias = bp
        buf = torch.empty((R, Vp), device=h.device, dtype=torch.float32)
        ops.gemm(h, arena.bf16(E), out=buf[:, :V], bias=bias)
        if bad_ids or top_k:
            ops.logits_filter(buf, V, bad_ids, top_k)
        nll, _ = ops.softmax_ce(buf, targets, V)                       # loss only; the gradient pass runs in backward with the weights
        ctx.saved = (h, buf, targets, head, arena, V, Vp)
        return -nll

    @staticmethod
    def backward(ctx, dlogp):
        h, buf, targets, head, arena, V, Vp = ctx.saved
        # d/dlogits of sum_r w_r * logp_r = -w_r * (softmax - onehot): the CE kernel writes (softmax - onehot) * grad_scale * row_weight
        w = (-dlogp).contiguous().float()
        ops.softmax_ce(buf, targets, V, dlogits=buf, row_weight=w)
        dl = ops.cast_bf16(buf)                                        # bf16 operand of the two LM-head gradient GEMMs
        E = head.decoder.weight
        ops.gemm(dl[:, :V], h, a_mn_major=True, b_mn_major=True, out=arena.grad(E), accumulate=True)
        ops.colsum(dl[:, :V], arena.grad(head.bias))
        dh = ops.gemm(dl[:, :V], arena.bf16(E), b_mn_major=True)
        return dh, None, None, None, None, None, None


def sequence_log_probs(decoder, sequences, encoder_hidden_states, encoder_attention_mask, bad_ids=(), top_k=0, targets=None):
    """sequences int64 [B, L] (BOS first) -> log-probabilities [B, L-1] of sequences[:, 1:] (or of `targets` [B, L-1]) under the
    decoder, teacher-forced on sequences[:, :-1], with the rollout's logit processors (bad-word removal, top-k) applied before the
    softmax; differentiable w.r.t. the decoder (and the encoder states)."""
    seq = sequences.cuda()
    B, L = seq.shape
    inp = seq[:, :-1].contiguous()
    x, _, T = decoder.hidden_states(inp, None, encoder_hidden_states, encoder_attention_mask)
    arena = get_arena(_root_of(decoder))
    targets = (seq[:, 1:] if targets is None else targets.cuda()).contiguous().view(-1)
    lp = _SeqLogProbFn.apply(decoder.head_input(x), decoder._core.embeddings.LayerNorm.weight, targets, decoder._head, arena, tuple(int(b) for b in bad_ids),
                             int(top_k or 0))
    return lp.view(B, L - 1)


def scst_loss(input, seq, reward_sampling, reward_greedy, scores_weights, pad_token_id):
    """vilmedic/blocks/rl/SCST.py:12-44 on [B, T] log-probabilities (tiny host-side tensor algebra over B*T numbers)."""
    input = torch.where(torch.isinf(input), torch.zeros_like(input), input)
    mask = (seq > pad_token_id).float()
    input = input * mask
    input = input / torch.sum(mask)
    delta_rewards = [torch.as_tensor(rs, dtype=torch.float32, device=input.device) - torch.as_tensor(rg, dtype=torch.float32, device=input.device)
                     for rs, rg in zip(reward_sampling, reward_greedy)]
    loss = [scores_weights[i] * (-input * r.unsqueeze(-1).expand_as(input)) for i, r in enumerate(delta_rewards)]
    loss = sum([torch.sum(l) for l in loss])
    delta_reward = torch.mean(torch.stack(delta_rewards))
    delta_reward_per_metric = torch.mean(torch.stack(delta_rewards), dim=-1)
    return loss, delta_reward, delta_reward_per_metric


class SCST(nn.Module):
    def __init__(self, decoder, dl, scores, scores_args=None, scores_weights=None, top_k=None, use_nll=False):
        super().__init__()
        dataset = dl.dataset
        if hasattr(dataset, "tokenizer"):
            self.tokenizer, self.max_length = dataset.tokenizer, dataset.tokenizer_max_len
        elif hasattr(dataset, "tgt_tokenizer"):
            self.tokenizer, self.max_length = dataset.tgt_tokenizer, dataset.tgt_tokenizer_max_len
        else:
            raise NotImplementedError("Where is tokenizer in dataset?")
        object.__setattr__(self, "decoder", decoder)          # not a sub-module: the model that owns the decoder registers it
        self.top_k = top_k
        self.use_nll = use_nll
        self.bos_token_id = decoder.config.bos_token_id
        self.eos_token_id = decoder.config.eos_token_id
        self.pad_token_id = decoder.config.pad_token_id
        assert scores is not None
        if not isinstance(scores, (list, tuple)):
            scores = [scores]
        self.scores = scores
        if len(scores) > 1 or use_nll:
            assert scores_weights is not None, "You need to mention scores_weights"
            assert isinstance(scores_weights, (list, tuple)), "scores_weights must be a list"
            assert len(scores_weights) == len(scores) + int(use_nll), "Mention (nll_weight +) as much scores_weights as scores"
            self.scores_weights = list(scores_weights)
        else:
            self.scores_weights = [1.0]
        if scores_args is not None:
            if not isinstance(scores_args, (list, tuple)):
                scores_args = [scores_args]
            assert len(scores_args) == len(scores), "You need to mention as much scores_args as scores"
        else:
            scores_args = [None] * len(scores)
        self.scores_args = scores_args
        self.scorers, self.scorers_index = [], []
        for score, args in zip(scores, scores_args):
            if callable(score):
                self.scorers.append(score)
                self.scorers_index.append(None)
                continue
            try:
                from vilmedic.blocks.scorers.scores import REWARD_COMPLIANT
            except Exception as e:
                raise NotImplementedError("scorer %r: the reference's scorers (CPU string metrics) are outside the hot path and not "
                                          "importable here; pass a callable scorer(refs, hyps) -> rewards" % (score,)) from e
            assert score.lower() in REWARD_COMPLIANT, "{} not in {}".format(score, list(REWARD_COMPLIANT.keys()))
            scorer, idx = REWARD_COMPLIANT[score.lower()]
            self.scorers.append(scorer(**args) if args is not None else scorer())
            self.scorers_index.append(idx)

    def forward_greedy(self, input_ids, encoder_hidden_states, encoder_attention_mask):
        assert not torch.is_grad_enabled(), "Please add torch.no_grad() decorator"
        batch_size = input_ids.shape[0]
        out = self.decoder.generate(input_ids=torch.ones((batch_size, 1), dtype=torch.long).cuda() * self.bos_token_id,
                                    max_length=self.max_length, num_beams=1, num_return_sequences=1, return_dict_in_generate=True,
                                    output_scores=True, encoder_hidden_states=encoder_hidden_states.detach(),
                                    encoder_attention_mask=encoder_attention_mask.detach(), forced_eos_token_id=True, use_cache=True)
        reward_greedy, hyp_list, ref_list = self.get_reward(out.sequences.detach(), input_ids)
        return reward_greedy, hyp_list, ref_list

    def forward_sampling(self, input_ids, attention_mask, encoder_hidden_states, encoder_attention_mask, reward_greedy):
        assert torch.is_grad_enabled()
        batch_size = input_ids.shape[0]
        if self.use_nll:
            nll_loss = self.decoder(input_ids=input_ids.cuda(), attention_mask=attention_mask.cuda(), encoder_hidden_states=encoder_hidden_states,
                                    encoder_attention_mask=encoder_attention_mask, labels=input_ids.cuda())["loss"]
        bad = [[self.pad_token_id], [self.bos_token_id]]
        with torch.no_grad():
            out = self.decoder.generate(input_ids=torch.ones((batch_size, 1), dtype=torch.long).cuda() * self.bos_token_id,
                                        max_length=self.max_length, num_beams=1, num_return_sequences=1,
                                        encoder_hidden_states=encoder_hidden_states.detach(),
                                        encoder_attention_mask=encoder_attention_mask.detach() if encoder_attention_mask is not None else None,
                                        bad_words_ids=bad, top_k=self.top_k, forced_eos_token_id=True, output_scores=True, do_sample=True,
                                        use_cache=True, return_dict_in_generate=True)
        sampled_ids = out.sequences[:, 1:].contiguous()
        reward_sampling, hyp_list, _ = self.get_reward(sampled_ids, input_ids)
        sampled_logits = sequence_log_probs(self.decoder, out.sequences, encoder_hidden_states, encoder_attention_mask,
                                            bad_ids=[b[0] for b in bad], top_k=self.top_k or 0)
        loss, delta_reward, delta_reward_per_metric = scst_loss(sampled_logits, sampled_ids, reward_sampling, reward_greedy,
                                                                self.scores_weights[-len(self.scores):], self.pad_token_id)
        if self.use_nll:
            loss = loss + self.scores_weights[0] * nll_loss
        return loss, delta_reward, delta_reward_per_metric, reward_sampling, hyp_list

    def get_reward(self, rollout_input_ids, input_ids):
        hyp_list, ref_list = [], []
        for h, r in zip(rollout_input_ids, input_ids):
            hyp_list.append(self.tokenizer.decode(h, skip_special_tokens=True, clean_up_tokenization_spaces=False))
            ref_list.append(self.tokenizer.decode(r, skip_special_tokens=True, clean_up_tokenization_spaces=False))
        reward = []
        for scorer, idx in zip(self.scorers, self.scorers_index):
            out = scorer(ref_list, hyp_list)
            reward.append(out if idx is None else out[idx])
        return reward, hyp_list, ref_list

    def __repr__(self):
        return "SCST\n" + json.dumps({"Scores": str(self.scores), "scores_args": str(self.scores_args),
                                      "scores_weights": str(self.scores_weights), "Generate": {"top_k": self.top_k}}, indent=4)
