"""Evaluator throughput of the C4Net stand-in (the CNN is outside the path; this only tells how the PyTorch side should be
driven): eager autocast vs bf16 weights vs channels_last vs CUDA-graph replay.  python tools/exp_cnn_throughput.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nets = importlib.import_module("alphazero-al_b200.nets")

torch.manual_seed(0)
dev = "cuda:0"
net = nets.C4Net(device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for B in ((16384,) if len(sys.argv) < 2 else tuple(int(x) for x in sys.argv[1].split(','))):
    planes = (torch.rand(B, 3, 6, 7, device=dev) > 0.6).float()
    planes[:, 1] *= 1 - planes[:, 0]
    mask = torch.ones(B, 7, dtype=torch.uint8, device=dev)
    net.max_batch = 32768
    ms = timeit(lambda: net.predict_device(planes, mask))
    print(f"B={B:6d} eager autocast max_batch=32768 : {ms:8.3f} ms  {B / ms / 1e3:8.3f} M evals/s", flush=True)
    ms = timeit(lambda: net.predict_device(planes, mask, autocast=False))
    print(f"B={B:6d} eager fp32 (tf32 off)          : {ms:8.3f} ms  {B / ms / 1e3:8.3f} M evals/s", flush=True)
    # bf16 weights, no autocast
    import copy
    nb = copy.deepcopy(net).to(torch.bfloat16)
    pb = planes.to(torch.bfloat16)
    with torch.no_grad():
        ms = timeit(lambda: nb(pb, mask))
    print(f"B={B:6d} bf16 weights, no autocast      : {ms:8.3f} ms  {B / ms / 1e3:8.3f} M evals/s", flush=True)
    ncl = copy.deepcopy(nb).to(memory_format=torch.channels_last)
    with torch.no_grad():
        ms = timeit(lambda: ncl(pb, mask))
    print(f"B={B:6d} bf16 weights, channels_last w  : {ms:8.3f} ms  {B / ms / 1e3:8.3f} M evals/s", flush=True)
    # CUDA graph of the autocast path and of the bf16 path
    for name, fn in (("graph autocast", lambda: net.predict_device(planes, mask)), ("graph bf16 weights", lambda: nb(pb, mask))):
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                out = fn()
            ms = timeit(g.replay)
            print(f"B={B:6d} {name:30s}: {ms:8.3f} ms  {B / ms / 1e3:8.3f} M evals/s", flush=True)
        except Exception as e:
            print(f"B={B:6d} {name}: failed {type(e).__name__}: {e}", flush=True)
    del nb, ncl

# which kernels dominate (one profiler pass, eager autocast, B=16384)
B = 16384
planes = (torch.rand(B, 3, 6, 7, device=dev) > 0.6).float()
mask = torch.ones(B, 7, dtype=torch.uint8, device=dev)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        net.predict_device(planes, mask)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
