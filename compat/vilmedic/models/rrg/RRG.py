from vilmedic_b200.models.rrg.RRG import RRG  # noqa: F401
