"""Mirror of vilmedic.models for the hot path: RRG (+RRG_HF), ConVIRT, MVQA compositions over the B200 blocks."""
from .rrg.RRG import RRG  # noqa: F401
