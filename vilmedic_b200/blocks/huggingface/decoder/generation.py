"""Greedy / beam-search decoding for the B200 decoder tower — placeholder until the KV-cache step
kernels land; `generate` re-runs the full prefix each step (correct, O(T^2)), selection logic follows HF
GenerationMixin._beam_search / the reference's ensemble loop (vilmedic/blocks/huggingface/decoder/beam_search.py:222-342)."""
import torch


class GenerationMixinB200:
    @torch.no_grad()
    def next_token_logits(self, input_ids, encoder_hidden_states=None, encoder_attention_mask=None):
        """fp32 logits [B, V] of the last position."""
        x, B, T = self.hidden_states(input_ids, None, encoder_hidden_states, encoder_attention_mask)
        last = x.view(B, T, -1)[:, -1].contiguous()
        return self.lm_logits(last)

    @torch.no_grad()
    def generate(self, input_ids=None, encoder_hidden_states=None, encoder_attention_mask=None, max_length=None,
                 num_beams=1, bos_token_id=None, eos_token_id=None, pad_token_id=None, length_penalty=1.0,
                 ensemble=None, **kwargs):
        from .beam import beam_search
        models = [self] if ensemble is None else list(ensemble)
        enc = encoder_hidden_states if isinstance(encoder_hidden_states, (list, tuple)) else [encoder_hidden_states] * len(models)
        msk = encoder_attention_mask if isinstance(encoder_attention_mask, (list, tuple)) else [encoder_attention_mask] * len(models)
        cfg = self.config
        return beam_search(models, enc, msk, input_ids=input_ids, max_length=max_length or 20, num_beams=num_beams,
                           bos_token_id=cfg.bos_token_id if bos_token_id is None else bos_token_id,
                           eos_token_id=cfg.eos_token_id if eos_token_id is None else eos_token_id,
                           pad_token_id=cfg.pad_token_id if pad_token_id is None else pad_token_id,
                           length_penalty=length_penalty)
