import torch, torch.nn.functional as F
torch.manual_seed(0)
x = torch.randn(4, 20, 12, 12); w = torch.randn(50, 20, 5, 5).sign() * 0.05; b = torch.randn(50)
ref = F.conv2d(x, w, b)
def err(): 
    y = F.conv2d(x.cuda(), w.cuda(), b.cuda()).cpu(); return float((y-ref).abs().max()/ref.abs().max())
print('default', torch.backends.cudnn.allow_tf32, err())
torch.backends.cudnn.allow_tf32 = False
print('allow_tf32=False', err())
try:
    torch.backends.cudnn.conv.fp32_precision = 'ieee'; print('conv.fp32_precision=ieee', err())
except Exception as e: print('conv.fp32_precision err', e)
try:
    torch.backends.fp32_precision = 'ieee'; print('backends.fp32_precision=ieee', err())
except Exception as e: print('fp32_precision err', e)
torch.backends.cudnn.enabled = False
print('cudnn disabled', err())
