"""Launch sequence for ncu: the fused stem-statistics kernel on the C2 re-rank batch (707 images of 256^2)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200.resnet import ResNetEncoder
enc = ResNetEncoder(seed=2000).to("cuda").eval()
x = torch.rand(707, 3, 256, 256, device="cuda")
for _ in range(3):
    enc.style_features(x)
torch.cuda.synchronize()
print("ok")
