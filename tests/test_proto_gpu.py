"""`proto: <local HuggingFace directory>` (vilmedic/blocks/huggingface/encoder/encoder_model.py:20-22, decoder/decoder_model.py:17-21 —
every shipped RRG / SELFSUP config sets one): BERT and RoBERTa checkpoints written by `save_pretrained` load into the kernel towers and
reproduce the HF modules (fp32, CPU) — hidden states, pooled output, LM loss and gradients — including RoBERTa's padding-aware position
ids, the token-type row and the dense -> GELU -> LayerNorm LM-head transform."""
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(vocab_size=400, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, max_position_embeddings=64,
          hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def _ids(pad, B=3, T=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, KW["vocab_size"], (B, T), generator=g)
    mask = torch.ones(B, T, dtype=torch.long)
    ids[1, 8:] = pad
    mask[1, 8:] = 0
    ids[2, 5:] = pad
    mask[2, 5:] = 0
    return ids, mask


def _rel(a, b):
    return ((a.float().cpu() - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("family", ["bert", "roberta"])
def test_encoder_proto_matches_hf(cuda_dev, tmp_path, family):
    from transformers import BertConfig, BertModel, RobertaConfig, RobertaModel
    from vilmedic_b200.blocks.huggingface.encoder.encoder_model import EncoderModel
    torch.manual_seed(0)
    hf = (BertModel(BertConfig(**KW)) if family == "bert" else RobertaModel(RobertaConfig(**KW))).eval()
    with torch.no_grad():
        hf.embeddings.token_type_embeddings.weight.normal_(0, 0.5)          # make the token-type row matter
    hf.config._attn_implementation = "eager"
    hf.save_pretrained(tmp_path)
    enc = EncoderModel({"proto": str(tmp_path), "add_pooling_layer": False}).cuda().train()
    ids, mask = _ids(hf.config.pad_token_id)
    want = hf(input_ids=ids, attention_mask=mask).last_hidden_state
    out = enc(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    valid = mask.bool()
    err = (out.float().cpu() - want)[valid].abs().max().item()
    assert err <= 3e-2 + 2 ** -7 * want.abs().max().item(), err
    w = torch.randn(want.shape, generator=torch.Generator().manual_seed(1)) * valid[..., None]
    (want * w).sum().backward()
    (out.float() * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    e, h = enc.encoder, hf
    assert _rel(e.embeddings.token_type_embeddings.weight.grad[0], h.embeddings.token_type_embeddings.weight.grad[0]) <= 6e-2
    assert _rel(e.embeddings.position_embeddings.weight.grad, h.embeddings.position_embeddings.weight.grad) <= 6e-2
    assert _rel(e.embeddings.word_embeddings.weight.grad, h.embeddings.word_embeddings.weight.grad) <= 6e-2
    assert _rel(e.encoder.layer[0].intermediate.dense.weight.grad, h.encoder.layer[0].intermediate.dense.weight.grad) <= 6e-2


@pytest.mark.parametrize("family", ["bert", "roberta"])
def test_decoder_proto_loss_and_generate(cuda_dev, tmp_path, family):
    from transformers import BertConfig, BertLMHeadModel, RobertaConfig, RobertaForCausalLM
    from vilmedic_b200.blocks.huggingface.decoder.decoder_model import DecoderModel
    torch.manual_seed(1)
    cfg = (BertConfig if family == "bert" else RobertaConfig)(is_decoder=True, add_cross_attention=True, **KW)
    hf = (BertLMHeadModel if family == "bert" else RobertaForCausalLM)(cfg).eval()
    hf.config._attn_implementation = "eager"
    with torch.no_grad():
        for p in hf.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    hf.save_pretrained(tmp_path)
    dm = DecoderModel({"proto": str(tmp_path)}).cuda().train()
    assert set(dm.decoder.state_dict()) == set(hf.state_dict())
    ids, mask = _ids(cfg.pad_token_id, seed=3)
    enc = torch.randn(3, 7, KW["hidden_size"], generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).float()
    emask = torch.ones(3, 7, dtype=torch.long)
    want = hf(input_ids=ids, attention_mask=mask, encoder_hidden_states=enc, encoder_attention_mask=emask, labels=ids, use_cache=False)
    out = dm(ids, mask, encoder_outputs=enc.cuda(), encoder_attention_mask=emask.cuda())
    assert abs(out["loss"].item() - want.loss.item()) <= 5e-3 * abs(want.loss.item()), (out["loss"].item(), want.loss.item())
    want.loss.backward()
    out["loss"].backward()
    torch.cuda.synchronize()
    d = dm.decoder
    if family == "roberta":
        pairs = [(d.lm_head.dense.weight, hf.lm_head.dense.weight), (d.lm_head.layer_norm.weight, hf.lm_head.layer_norm.weight),
                 (d.lm_head.bias, hf.lm_head.bias), (d.roberta.embeddings.word_embeddings.weight, hf.roberta.embeddings.word_embeddings.weight)]
    else:
        t, ht = d.cls.predictions, hf.cls.predictions
        pairs = [(t.transform.dense.weight, ht.transform.dense.weight), (t.transform.LayerNorm.weight, ht.transform.LayerNorm.weight),
                 (t.bias, ht.bias), (d.bert.embeddings.word_embeddings.weight, hf.bert.embeddings.word_embeddings.weight)]
    for mine_p, hf_p in pairs:
        assert _rel(mine_p.grad, hf_p.grad) <= 6e-2
    # the head transform is part of the decode step too
    dm.eval()
    seq = dm.generate(input_ids=torch.full((3, 1), 2, dtype=torch.long, device="cuda"), encoder_hidden_states=enc.cuda(),
                      encoder_attention_mask=emask.cuda(), max_length=6, num_beams=1, bos_token_id=2, eos_token_id=3, pad_token_id=cfg.pad_token_id)
    with torch.no_grad():
        lg = hf(input_ids=seq[:, :1].cpu(), encoder_hidden_states=enc, encoder_attention_mask=emask, use_cache=False).logits[:, -1]
    top2 = torch.topk(lg, 2).values
    safe = (top2[:, 0] - top2[:, 1]) > 0.05 * lg.abs().max()
    assert torch.equal(seq[:, 1].cpu()[safe], lg.argmax(-1)[safe])
