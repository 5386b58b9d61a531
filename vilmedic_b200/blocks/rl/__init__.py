from .scst import SCST, scst_loss, sequence_log_probs  # noqa: F401
