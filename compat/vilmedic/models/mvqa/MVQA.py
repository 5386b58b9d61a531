from vilmedic_b200.models.mvqa.MVQA import MVQA  # noqa: F401
