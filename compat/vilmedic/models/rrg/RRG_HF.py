from vilmedic_b200.models.rrg.RRG_HF import RRG_HF  # noqa: F401
