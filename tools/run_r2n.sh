#!/bin/bash
# round 2, dedicated square: device check of the new carry chains, parity of the kernels that use it, quick numbers
python -m pytest tests/test_gpu_field.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import time, torch
from verkle_kzg_b200 import Engine
eng = Engine(0)
n = 148 * 2048 * 4
g = torch.Generator().manual_seed(1)
raw = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g); raw[:, 31] &= 0x1F
for mode in (1, 0, 2, 3, 4):
    a = raw.cuda()
    eng.probe_fq_sqr_dev(a, n, 200, mode); eng.sync()
    t = time.perf_counter(); eng.probe_fq_sqr_dev(a, n, 2000, mode); eng.sync(); dt = time.perf_counter() - t
    print("mode", mode, "(0 = dedicated square, 1 = general product a*a, 2 = Karatsuba a*a, 3 = general a*b, 4 = Karatsuba a*b):", "%.2f G products/s" % (n * 2000 / dt / 1e9))
eng.close()
PY
