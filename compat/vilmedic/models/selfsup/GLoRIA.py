from vilmedic_b200.models.selfsup.GLoRIA import GLoRIA  # noqa: F401
