"""One hot-path step (the bench's models, inputs and calls) for `ncu --profile-from-start off`.

The capture windows are chosen with MER_CUPROF="klass:first:count,..." (runtime.cu): launches [first, first + count) of a
kernel class are bracketed with cudaProfilerStart / Stop inside the library, so one profiled process captures a few
launches of every kernel of the step.  Class ids: 0 TF32 GEMM, 1 BF16X3 GEMM, 2 F16 GEMM, 10 attention_f16,
11 attention_tc, 12 LayerNorm, 13 positional conv, 14 conv0 (stats + apply), 15 long-key attention.
The fusion step's two kernels: scripts/bench_fusion_step.py --profile."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--long-audio", action="store_true", help="also one HuBERT forward on 8 x 10 s rows (499 frames)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
models = bench.build_models(dev)
dev_in = [x.to(dev) for x in bench.make_inputs(0, a.clips)]
bench.device_step(models, dev_in, a.clips, 1)  # MER_CUPROF windows count launches from the first one on
torch.cuda.synchronize()
if a.long_audio:
    wave = torch.randn(8, 160000, device=dev) * (3000.0 / 32768.0)
    models[1].forward(wave, normalize=True)
    torch.cuda.synchronize()
print("profile_step done", flush=True)
