"""One eager train step at batch 32, 256x256 (what bench.py times) -- the target of the ncu captures: ncu -k regex:<kernel> ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200.models import FaceVAE
from face_vae_b200.trainer import VAETrainer

torch.manual_seed(0)
m = FaceVAE().cuda().train()
tr = VAETrainer(m, use_cuda_graph=False)
x = torch.rand((32, 3, 256, 256), device="cuda")
eps = torch.randn((32, 4096), device="cuda")
for _ in range(int(os.environ.get("STEPS", "2"))):
    tr.step(x, eps)
torch.cuda.synchronize()
print("ok")
