"""Mirror of vilmedic.models for the hot path: RRG, RRG_HF, ConVIRT, GLoRIA, MVQA compositions over the B200 blocks."""
from .mvqa.MVQA import MVQA  # noqa: F401
from .rrg.RRG import RRG  # noqa: F401
from .rrg.RRG_HF import RRG_HF  # noqa: F401
from .selfsup.conVIRT import ConVIRT  # noqa: F401
from .selfsup.GLoRIA import GLoRIA  # noqa: F401
