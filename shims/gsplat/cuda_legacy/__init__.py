"""``gsplat.cuda_legacy``: the two pure helpers the reference imports from it."""
