"""vilmedic.blocks (hot-path sub-packages only) -> vilmedic_b200.blocks."""
