"""Dev tool: run ONE conv / gemm launch repeatedly (for `ncu --set full -k regex:... -s 3 -c 1`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scripts.bench_gemm as bg  # noqa: E402

kind = sys.argv[1]
args = tuple(int(v) for v in sys.argv[2:])
print(kind, args, bg.run(kind, args, iters=3))
