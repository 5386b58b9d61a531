from vilmedic_b200.models import MVQA, RRG, RRG_HF, ConVIRT, GLoRIA  # noqa: F401
