import os, sys
sys.path.insert(0, os.getcwd())
import torch
from eavsr_b200 import ops
import eavsr_b200.model as M
dev = torch.device("cuda:0")
blk = M._AdaptBlockOffset(64, 8).to(dev, torch.bfloat16)
mk = lambda: torch.randn(1, 64, 272, 480, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
a, b = mk(), mk()
with torch.no_grad():
    for _ in range(3):
        blk._mix(a, b)
torch.cuda.synchronize()
