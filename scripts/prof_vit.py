"""Launch sequence for ncu: one CLIP ViT-L/14 encode call of 500 uint8 images through drag_vit_encode (C2's batch)."""
import sys, torch
sys.path.insert(0, '.')
from domain_rag_b200 import clip
model, _ = clip.load("ViT-L/14", device="cuda", seed=2000)
model.visual.fold_layernorm(len(sys.argv) > 1 and sys.argv[1] == "fold")
x = torch.randint(0, 255, (500, 3, 224, 224), dtype=torch.uint8, device="cuda")
for _ in range(2):
    model.encode_image(x, normalize=True)
torch.cuda.synchronize()
print("ok")
