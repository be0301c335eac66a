"""Context network for the end-to-end train-throughput figure (bench.py --workload train-*): a torchvision ResNet
backbone (stride 32, as the reference's ResNetBackbone wraps it) and a DeepLabV3+-style depthwise-separable ASPP
head with a projection branch.  Stock cuDNN / ATen modules only: this is the surrounding workload the hierarchical
loss runs in, not part of the product (backbone and head are out of scope)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _cbr(cin, cout, k=1, **kw):
    return nn.Sequential(nn.Conv2d(cin, cout, k, bias=False, **kw), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _sep(cin, cout, dilation=1):
    return nn.Sequential(_cbr(cin, cin, 3, padding=dilation, dilation=dilation, groups=cin), _cbr(cin, cout, 1))


class ContextSegNet(nn.Module):
    def __init__(self, depth: int, num_classes: int, proj_dim: int = 256, aspp: int = 512, c1: int = 48):
        super().__init__()
        import torchvision
        base = getattr(torchvision.models, f"resnet{depth}")(weights=None)
        self.stem = nn.Sequential(base.conv1, base.bn1, base.relu, base.maxpool)
        self.layers = nn.ModuleList([base.layer1, base.layer2, base.layer3, base.layer4])
        c_low, c_top = (64, 512) if depth in (18, 34) else (256, 2048)
        self.proj = nn.Sequential(_cbr(c_top, c_top, 1), nn.Conv2d(c_top, proj_dim, 1, bias=False))
        self.branches = nn.ModuleList([_cbr(c_top, aspp, 1)] + [_sep(c_top, aspp, d) for d in (12, 24, 36)])
        self.pool = _cbr(c_top, aspp, 1)
        self.fuse = _cbr(5 * aspp, aspp, 1)
        self.low = _cbr(c_low, c1, 1)
        self.refine = nn.Sequential(_sep(aspp + c1, aspp), _sep(aspp, aspp))
        self.cls = nn.Conv2d(aspp, num_classes, 1)

    def forward(self, x):
        x = self.stem(x)
        feats = []
        for layer in self.layers:
            x = layer(x)
            feats.append(x)
        top = feats[-1]
        emb = F.normalize(self.proj(top), dim=1)                              # [B, proj_dim, H/32, W/32]
        pooled = self.pool(F.adaptive_avg_pool2d(top, 1)).expand(-1, -1, *top.shape[2:])
        y = self.fuse(torch.cat([b(top) for b in self.branches] + [pooled], dim=1))
        low = self.low(feats[0])
        y = F.interpolate(y, size=low.shape[2:], mode="bilinear", align_corners=False)
        logits = self.cls(self.refine(torch.cat([y, low], dim=1)))            # [B, C, H/4, W/4]
        return logits, emb, feats[2]                                          # c3 (stride 16) feeds the aux head
