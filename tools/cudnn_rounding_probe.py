"""GPU box only: where does the reference's GPU path (cuDNN half conv2d called by aten) round when a convolution has a
bias — once after conv + bias, or after the convolution and again after the bias add?  (genUpsampleBlock, models.py:29-30.)
Decides which of the two the engine's EPI_BIAS_PRELU epilogue mimics.  Prints mismatch counts against both candidates."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
for cout in (256, 576):
  x = torch.rand(3, 64, 40, 56, generator=g).half().cuda()
  w = (torch.randn(cout, 64, 3, 3, generator=g) * 0.05).half().cuda()
  b = (torch.randn(cout, generator=g) * 0.1).half().cuda()
  y = F.conv2d(x, w, b, padding=1)
  y32 = F.conv2d(x.double(), w.double(), None, padding=1)
  fused = (y32 + b.double().view(1, -1, 1, 1)).half()
  sep = (y32.half().double() + b.double().view(1, -1, 1, 1)).half()
  n = y.numel()
  print('cout %d: cuDNN half conv+bias differs from q(conv+bias) in %.4f %% of the outputs, from q(q(conv)+bias) in %.4f %%'
        % (cout, 100.0 * (y != fused).sum().item() / n, 100.0 * (y != sep).sum().item() / n))
  y0 = F.conv2d(x, w, None, padding=1)
  print('          without bias: differs from q(conv) in %.4f %%' % (100.0 * (y0 != y32.half()).sum().item() / n))
# PReLU / scale / add on half tensors: fp32 opmath, one rounding each
v = torch.randn(100000, generator=g).half().cuda()
a = torch.tensor([0.1003], device='cuda').half()
print('prelu mismatches', (F.prelu(v, a) != torch.where(v >= 0, v.float(), v.float() * a.float()).half()).sum().item())
s = torch.tensor([0.2517], device='cuda').half()
print('scale mismatches', ((v * s) != (v.float() * s.float()).half()).sum().item())
