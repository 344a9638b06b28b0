"""Developer tool: times one fused-linear launch shape on the three kernels (fp32 SIMT is skipped for big shapes).
    MESM_DEBUG_LINEAR_ITERS=20 python tools/linear_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MESM_DEBUG_LINEAR_ITERS", "20")
from tests.test_gpu_linear import _run  # noqa: E402

for case in (dict(M=598276, N=256, K=256), dict(M=598276, N=256, K=256, res=True, ln=True), dict(M=602372, N=512, K=256),
             dict(M=240000, N=256, K=2818, fold=True, act=1, ln=True)):
    for mode in (1, 2, 3) + ((4,) if case["K"] > 2000 else ()):
        _run(use_tc=mode, f16_inputs=mode == 4, **case)
