"""GPU triage: per-stage error of the depth engine vs the fp32 oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from oracle import depth as OD
from visiondepth3d_b200.depth_engine import DepthEngine
from visiondepth3d_b200.depth_weights import CONFIGS, hf_config
from transformers import DepthAnythingForDepthEstimation

name, h, w = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("vits", 70, 98)
torch.manual_seed(0)
sd = DepthAnythingForDepthEstimation(hf_config(name)).eval().state_dict()
cfg = CONFIGS[name]
e = DepthEngine(name, h, w)
# GEMM sanity
rng = np.random.default_rng(0)
for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 256), (300, 200, 72)]:
    A = (rng.standard_normal((M, K)) * .5).astype(np.float16); B = (rng.standard_normal((N, K)) * .5).astype(np.float16)
    ref = A.astype(np.float32) @ B.astype(np.float32).T
    out = e.gemm(A, B)
    print("gemm", (M, N, K), "maxerr", np.abs(out - ref).max(), "refmax", np.abs(ref).max(), flush=True)
e.load_state_dict(sd)
torch.manual_seed(2)
px = torch.randn(3, h, w)
with torch.no_grad():
    ref, parts = OD.forward(sd, cfg, px, return_parts=True)
out = e.forward(px.numpy())
D = cfg["hidden"]; ph, pw = h // 14, w // 14; NT = ph * pw + 1
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
x = e.get_buffer("x", (NT, D), np.float32)
print("x(final residual) rel err", rel(x, parts["x"][0].numpy()))
for i in range(4):
    t = e.get_buffer(f"tap{i}", (ph * pw, D), np.float16).astype(np.float32)
    print(f"tap{i} rel err", rel(t, parts["taps"][i][0, 1:].numpy()))
for i in range(4):
    f = parts["feats"][i][0].permute(1, 2, 0).numpy()
    g = e.get_buffer(f"f{i}", f.shape, np.float16).astype(np.float32)
    print(f"feat{i} {f.shape} rel err", rel(g, f))
for j in range(4):
    f = parts["fused"][j][0].permute(1, 2, 0).numpy()
    g = e.get_buffer(f"fused{j}", f.shape, np.float16).astype(np.float32)
    print(f"fused{j} {f.shape} rel err", rel(g, f))
r = ref.numpy()
print("depth rel-to-range err", float(np.abs(out - r).max() / (r.max() - r.min())), "range", r.min(), r.max())
