"""evaluation(models, config, dl) for classification heads — mirror of vilmedic/blocks/classifier/evaluation.py:7-64 (the
`eval_func` of MVQA, vilmedic/models/mvqa/MVQA.py:38): every model scores every batch, logits are averaged over models,
the loss is the mean over batches and models.  Host-side bookkeeping only; the model forwards run on the kernels.
`attentions` are never materialised by the fused attention kernels, so that optional post-processing entry is absent
(the reference only fills it when a model returns the key)."""
import numpy as np
import torch


def evaluation(models, config, dl, **kwargs):
    logits = np.array([])
    labels = np.array([])
    losses = np.array([])
    cumulative_index = 0
    with torch.no_grad():
        for num_batch, batch in enumerate(dl):
            label = batch["labels"]
            batch_size = label.shape[0]
            num_classes = label.shape[1] if label.dim() > 1 else None
            batch = {k: v.cuda() if (isinstance(v, torch.Tensor) and torch.cuda.is_available()) else v for k, v in batch.items()}
            results = [model(**batch) for model in models]
            if num_batch == 0:                      # pre-allocate (reference :27-36)
                n_out = results[0]["output"].shape[-1]
                logits = np.zeros((len(dl.dataset), len(models), n_out))
                labels = np.zeros((len(dl.dataset), num_classes)) if num_classes is not None else np.zeros((len(dl.dataset),))
                losses = np.zeros((len(dl), len(models)))
            for j, r in enumerate(results):
                logits[cumulative_index:cumulative_index + batch_size, j] = r["output"].float().cpu().numpy()
                losses[num_batch][j] = r["loss"].cpu().item()
            labels[cumulative_index:cumulative_index + batch_size] = label.cpu().numpy()
            cumulative_index += batch_size
    preds = np.mean(logits, axis=1)
    loss = np.mean(losses)
    return {"loss": loss, "refs": labels, "hyps": preds, "logits": logits}
