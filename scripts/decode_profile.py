import torch, sys
sys.path.insert(0, '/root/repo')
import seghiero_b200 as sb
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(8, 28, 2048, 2048, generator=g, device='cuda', dtype=torch.float32).bfloat16()
for _ in range(3):
    p, _ = sb.hierarchical_argmax(x, [19, 7, 2])
torch.cuda.synchronize()
