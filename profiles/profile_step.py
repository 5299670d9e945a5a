"""One forward step of the benchmark workload between cudaProfilerStart/Stop, for Nsight Compute:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:solve_v1_kernel -c 1 -o gpurun_out/solve python profiles/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ml_quant_b200 import runtime  # noqa: E402

batch = int(os.environ.get('LSQ_PROFILE_BATCH', '512'))
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
model = runtime.build_model('imagenet_resnet18_ls1w_ls2a', dev)
runtime.calibrate(model, (3, 224, 224))
if os.environ.get('LSQ_PROFILE_NOFUSE') != '1':
    runtime.optimize_for_inference(model)
x = torch.randn(batch, 3, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled one step, batch', batch)
