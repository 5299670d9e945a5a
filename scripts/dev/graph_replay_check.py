"""Development aid: replay time of CUDA graphs captured one after the other (are later captures slower?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_quant_b200 import runtime
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
model = runtime.build_model('imagenet_resnet18_ls1w_ls2a', dev)
runtime.calibrate(model, (3, 224, 224))
runtime.optimize_for_inference(model)
B = 512
x = torch.randn(B, 3, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(3): model(x)
def t_replay(g, n=10):
    g(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gs = []
for i in range(3):
    gs.append(runtime.GraphedForward(model, x))
    print('after capture', i, 'replay ms of each graph:', [round(t_replay(g), 2) for g in gs], 'mem GB', round(torch.cuda.memory_reserved() / 2**30, 1), flush=True)
