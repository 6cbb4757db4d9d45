"""one moe_conv3x3_c64 call (for compute-sanitizer / ncu):  python tools/one_conv.py n h w r epi"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tools.gpu_diag as D
from moephoto_b200 import imageProcess as IP
n, h, w, r, epi = [int(a) for a in sys.argv[1:6]]
eng = IP.getEngine(0)
g = torch.Generator().manual_seed(0)
x = (torch.randn(n, h, w, 64, generator=g) * 0.5).half().cuda()
wt = (torch.randn(64 * r * r, 64, 3, 3, generator=g) * 0.05).half()
bias = (torch.randn(64 * r * r, generator=g) * 0.1).half() if epi == 3 else None
skip = torch.randn(n, h, w, 64, generator=g).half().cuda() if epi == 2 else None
ref = D.conv_ref(x, wt.cuda(), None if bias is None else bias.cuda(), r, epi, 0.25, skip)
out = D.run_conv(eng, x, wt, bias, r, epi, 0.25, skip).float()
print('max|d|', (out - ref).abs().max().item(), 'nan', torch.isnan(out).sum().item())
