import sys, os
sys.path.insert(0, "all-in-one-deflicker_b200")
import torch
from b200 import nn as K
g = torch.Generator().manual_seed(0)
f1 = torch.randn(1, 256, 135, 240, generator=g).cuda(); f2 = torch.randn(1, 256, 135, 240, generator=g).cuda()
for _ in range(3):
    p = K.corr_build(f1, f2); torch.cuda.synchronize(); del p
