"""evaluation(models, config, dl) of RRG_HF — mirror of
vilmedic/blocks/huggingface/encoder_decoder/vision_multi_evaluation.py:16-129 (`eval_func` of vilmedic/models/rrg/RRG_HF.py:104).

Per batch: 5-D images are flattened to B*N crops, encoded once, concatenated to [B, N*S, D] (+ enc_to_dec_proj), the image-presence
mask is expanded to a patch mask and handed to `generate` together with the precomputed encoder states (:62-104); 4-D images go
through `generate(images, ...)` (:107-111).  Generation arguments are assembled into a GenerationConfig-like object exactly as
the reference does (:42-57: max_length = tokenizer_max_len, num_beams = config.beam_width, length_penalty if configured,
decoder_start_token_id = [CLS]); hyps / refs are decoded with the dataset tokenizer (:117-127).
Like the reference, only models[0] decodes (no ensembling on this path)."""
from types import SimpleNamespace

import torch
import torch.nn as nn


def get_special_token_ids(model, tokenizer):
    bos_token_id = model.config.bos_token_id
    eos_token_id = model.config.eos_token_id
    pad_token_id = model.config.pad_token_id
    if None in [bos_token_id, eos_token_id, pad_token_id]:
        bos_token_id = tokenizer.vocab[tokenizer.cls_token]
        eos_token_id = tokenizer.vocab[tokenizer.sep_token]
        pad_token_id = tokenizer.vocab[tokenizer.pad_token]
    return bos_token_id, eos_token_id, pad_token_id


def evaluation(models, config, dl, **kwargs):
    first_model = models[0]
    if isinstance(first_model, nn.DataParallel):
        first_model = first_model.module
    hf_model = first_model.model if hasattr(first_model, "model") else first_model
    try:
        ref_str = "input_ids"
        tokenizer = dl.dataset.tokenizer
        max_len = dl.dataset.tokenizer_max_len
    except AttributeError:
        ref_str = "decoder_input_ids"
        tokenizer = dl.dataset.tgt_tokenizer
        max_len = dl.dataset.tgt_tokenizer_max_len
    refs, hyps = [], []
    bos_id, eos_id, pad_id = get_special_token_ids(hf_model, tokenizer)
    gen_args = {"bos_token_id": bos_id, "eos_token_id": eos_id, "pad_token_id": pad_id, "num_return_sequences": 1,
                "max_length": max_len, "use_cache": True}
    if getattr(config, "length_penalty", None) is not None:
        gen_args["length_penalty"] = config.length_penalty
    if getattr(config, "beam_width", None) is not None:
        gen_args["num_beams"] = config.beam_width
    gen_conf = SimpleNamespace(**gen_args, decoder_start_token_id=tokenizer.cls_token_id)
    with torch.no_grad():
        for batch in dl:
            batch = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            images = batch["images"]
            img_mask = batch.get("images_mask", None)
            if images.dim() == 5:
                B, N, C, H, W = images.shape
                if img_mask is None:
                    img_mask = torch.ones((B, N), dtype=torch.bool, device=images.device)
                else:
                    img_mask = img_mask.to(images.device).bool()
                flat_hidden = hf_model.encoder(images.view(B * N, C, H, W))          # [B*N, S, D]
                S, D = flat_hidden.size(1), flat_hidden.size(2)
                concat_hidden = flat_hidden.reshape(B, N * S, D)
                if getattr(hf_model, "enc_to_dec_proj", None) is not None:
                    concat_hidden = hf_model.project(concat_hidden)
                attn_mask = img_mask.unsqueeze(-1).expand(B, N, S).reshape(B, N * S).long()
                out_ids = hf_model.generate(generation_config=gen_conf,
                                            encoder_outputs=SimpleNamespace(last_hidden_state=concat_hidden),
                                            attention_mask=attn_mask)
            elif images.dim() == 4:
                out_ids = hf_model.generate(images, generation_config=gen_conf)
            else:
                raise NotImplementedError(f"Unexpected images.dim() = {images.dim()}")
            for pred_ids, ref_ids in zip(out_ids, batch[ref_str]):
                hyps.append(tokenizer.decode(pred_ids, skip_special_tokens=True, clean_up_tokenization_spaces=False))
                refs.append(tokenizer.decode(ref_ids, skip_special_tokens=True, clean_up_tokenization_spaces=False))
    return {"refs": refs, "hyps": hyps}
