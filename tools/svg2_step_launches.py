"""One warm SVG2 step (HunyuanVideo-720p shape) for an ncu launch list: which kernels a sparse_core call launches."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from svgb200.models import hyvideo as hy  # noqa: E402

dev = torch.device("cuda:0")
H, S, D = 24, bench.S, bench.D
sap = hy.HunyuanSAPCore(bench.CTX, bench.F, bench.P, num_q_centroids=400, num_k_centroids=1000, top_p_kmeans=0.9,
                        min_kc_ratio=0.1, kmeans_iter_init=2, kmeans_iter_step=2, prompt_length=bench.PROMPT_LEN)
g = torch.Generator(device=dev).manual_seed(11)
q, k, v = (torch.randn(1, H, S, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
sap.sparse_core(q, k, v)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sap.sparse_core(q, k, v)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
