"""Run on the GPU box (gpurun): records what torch.rand / torch.randint produce on CUDA for known
(seed, offset) states, as the golden vector that pins the oracle's Philox restatement
(tests/test_oracle_c.py). Output: gpurun_out/cuda_philox.json -> copy to tests/golden/."""
import json
import os

import torch

dev = torch.device("cuda:0")
props = torch.cuda.get_device_properties(dev)
torch.zeros(1, device=dev)
gen = torch.cuda.default_generators[0]
cases = []
for seed, n, n_int, rng, pre in [(1, 512 * 20, 512, 256, 0), (77, 64 * 8, 64, 32, 3), (2 ** 40 + 5, 4096 * 20, 8192, 256, 1),
                                 (9, 400000, 300000, 1000, 0)]:
    torch.manual_seed(seed)
    for _ in range(pre):
        torch.rand(5, device=dev)
    off = gen.get_offset()
    r = torch.rand(n, device=dev)
    off1 = gen.get_offset()
    ints = torch.randint(0, rng, (n_int,), device=dev)
    off2 = gen.get_offset()
    idx = sorted(set([0, 1, 2, 3, 255, 256, 257, n // 2, n - 2, n - 1]))
    iidx = sorted(set([0, 1, 2, 3, n_int // 2, n_int - 1]))
    cases.append({"seed": seed, "offset": off, "n": n, "offset_after_rand": off1, "n_int": n_int, "range": rng,
                  "offset_after_int": off2, "rand_idx": idx, "rand_hex": [float(r[i]).hex() for i in idx],
                  "int_idx": iidx, "int_vals": [int(ints[i]) for i in iidx]})
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"torch": torch.__version__, "gpu": props.name, "sms": props.multi_processor_count,
           "max_threads_per_sm": props.max_threads_per_multi_processor, "cases": cases},
          open("gpurun_out/cuda_philox.json", "w"), indent=1)
print("ok")
