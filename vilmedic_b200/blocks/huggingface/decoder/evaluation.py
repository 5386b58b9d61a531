"""evaluation(models, config, dl) — mirror of vilmedic/blocks/huggingface/decoder/evaluation.py:20-85: encode with every
model, beam-search with `config.beam_width` / `config.length_penalty`, decode hyps/refs with the dataset tokenizer.

Reference defect resolved (SURVEY.md §8 defects #1): the reference decodes with model 0's decoder on the LAST model's
encoder output and ignores the rest (:64-78); the intended ensemble (sum of next-token logits over all models,
beam_search.py:254) is what runs here when len(models) > 1.
"""
import torch
import torch.nn as nn

from ....cfgutil import cfg_get


def get_special_token_ids(model, tokenizer):
    bos, eos, pad = model.config.bos_token_id, model.config.eos_token_id, model.config.pad_token_id
    if None in [bos, eos, pad]:
        bos = tokenizer.vocab[tokenizer.cls_token]
        eos = tokenizer.vocab[tokenizer.sep_token]
        pad = tokenizer.vocab[tokenizer.pad_token]
    return bos, eos, pad


def evaluation(models, config, dl, **kwargs):
    models = [m if not isinstance(m, nn.DataParallel) else m.module for m in models]
    hf_models = [m.dec.decoder for m in models]
    try:
        ref_str = "input_ids"
        tokenizer = dl.dataset.tokenizer
        max_len = dl.dataset.tokenizer_max_len
    except AttributeError:
        ref_str = "decoder_input_ids"
        tokenizer = dl.dataset.tgt_tokenizer
        max_len = dl.dataset.tgt_tokenizer_max_len
    bos, eos, pad = get_special_token_ids(hf_models[0], tokenizer)
    length_penalty = cfg_get(config, "length_penalty", None)
    beam_width = cfg_get(config, "beam_width", None)
    ref_list, hyp_list = [], []
    with torch.no_grad():
        for batch in dl:
            batch = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            bs = batch[ref_str].shape[0]
            encs, masks = [], []
            for m in models:
                e, em = m.encode(**batch)
                encs.append(e)
                masks.append(em)
            hyps = hf_models[0].generate(
                input_ids=torch.ones((bs, 1), dtype=torch.long, device="cuda") * bos,
                encoder_hidden_states=encs, encoder_attention_mask=masks, ensemble=hf_models,
                max_length=max_len, num_beams=beam_width or 1, bos_token_id=bos, eos_token_id=eos, pad_token_id=pad,
                length_penalty=1.0 if length_penalty is None else length_penalty)
            for h, r in zip(hyps, batch[ref_str]):
                hyp_list.append(tokenizer.decode(h, skip_special_tokens=True, clean_up_tokenization_spaces=False))
                ref_list.append(tokenizer.decode(r, skip_special_tokens=True, clean_up_tokenization_spaces=False))
    return {"refs": ref_list, "hyps": hyp_list}
