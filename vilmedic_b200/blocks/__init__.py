"""Mirror of vilmedic.blocks for the hot path: same class names, constructor kwargs and call
signatures as the reference, arithmetic on the sm_100a kernels."""
