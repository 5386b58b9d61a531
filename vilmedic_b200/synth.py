"""Seeded synthetic inputs of the reference batch shape and the BASELINE model shapes; used by bench.py and the tests."""
import torch

BOS, PAD, EOS = 0, 1, 2   # vilmedic/datasets/base/utils.py:24-25 order [CLS],[PAD],[SEP],[UNK],[MASK]; config/RRG/baseline-mimic.yml:15-16,28


def rrg_batch(batch, seq_len, vocab, image_size=224, seed=1234, n_images=None):
    """{'input_ids','attention_mask','images','images_mask'} as ImSeq's collate produces them (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    shape = (batch, 3, image_size, image_size) if n_images is None else (batch, n_images, 3, image_size, image_size)
    images = torch.randn(shape, generator=g)
    lens = torch.randint(seq_len // 2, seq_len + 1, (batch,), generator=g)
    ids = torch.randint(5, vocab, (batch, seq_len), generator=g)
    ids[:, 0] = BOS
    mask = torch.zeros(batch, seq_len, dtype=torch.long)
    for b in range(batch):
        L = int(lens[b])
        ids[b, L - 1] = EOS
        ids[b, L:] = PAD
        mask[b, :L] = 1
    return {"input_ids": ids, "attention_mask": mask, "images": images, "images_mask": None}


def vit_b16():
    return dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224,
                patch_size=16, num_channels=3)


def bert_base_decoder(vocab=30522, layers=12, dropout=0.1):
    return dict(vocab_size=vocab, hidden_size=768, num_hidden_layers=layers, num_attention_heads=12, intermediate_size=3072,
                max_position_embeddings=512, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout,
                layer_norm_eps=1e-12, bos_token_id=BOS, pad_token_id=PAD, eos_token_id=EOS)
