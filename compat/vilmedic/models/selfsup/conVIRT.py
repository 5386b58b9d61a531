from vilmedic_b200.models.selfsup.conVIRT import ConVIRT  # noqa: F401
