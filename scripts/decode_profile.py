"""Small decode workload for `ncu --set full` captures: full-resolution decode (k_decode_vec) and the fused decode from
the head's H/4 logits (k_decode_up4), bf16, 8 x 1024 x 2048 output pixels, uint8 labels and predictions."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seghiero_b200 as sb

g = torch.Generator(device='cuda').manual_seed(1)
b, h, w = 8, 1024, 2048
lab = torch.randint(0, 19, (b, h, w), generator=g, device='cuda').to(torch.uint8)
x = torch.randn(b, 28, h, w, generator=g, device='cuda', dtype=torch.float32).bfloat16()
xl = torch.randn(b, 28, h // 4, w // 4, generator=g, device='cuda', dtype=torch.float32).bfloat16()
for _ in range(3):
    sb.hierarchical_argmax(x, [19, 7, 2], label=lab, out_dtype=torch.uint8)
    sb.hierarchical_argmax(xl, [19, 7, 2], label=lab, size=(h, w), out_dtype=torch.uint8)
torch.cuda.synchronize()
