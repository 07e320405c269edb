"""`CorrBlock` with the reference's interface (src/models/stage_1/core/corr.py:16-54), on
b200_corr_build / b200_corr_lookup.  (The reference's optional native hook for this operator,
`alt_cuda_corr.forward`, corr.py:86-91, is what these two entry points stand in for.)"""
from b200 import nn as K


class CorrBlock:
    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, out=None):
        if num_levels != 4:
            raise NotImplementedError("the RAFT configuration of the reference uses 4 levels")
        self.num_levels, self.radius = num_levels, radius
        # `out`: a caller-owned pyramid buffer (RAFT keeps one per geometry so that its captured refinement graph
        # always reads the same addresses)
        self.pyramid = K.corr_build(fmap1.float().contiguous(), fmap2.float().contiguous(), out=out)

    def __call__(self, coords):
        return K.corr_lookup(self.pyramid, coords.float().contiguous(), self.radius)
