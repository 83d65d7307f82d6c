"""TEST INFRASTRUCTURE -- writes tests/golden/long_s1_heun250.pt: one full-length trajectory of the CPU oracle.

BASELINE.json asks for the final-latent error to be STATED; the short-grid fixtures say nothing about where the bf16
engine ends up after 498 evaluations.  This runs the oracle (fp32, CPU; pinned to the reference by
tests/test_oracle_golden.py) on LightningDiT-S/1, n=4 (8 model rows), 250-point Heun + CFG (scale 10 on 3 channels,
timestep_shift 0.3 -- the sampler settings of the benchmark) and stores the kept rows of the final state plus the
guided velocity of a few evaluations along the way.  ~3-4 minutes on 8 cores; run once, commit the output:

    python -m oracle.gen_golden_long
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")

import torch

from oracle import lightningdit_oracle as O

KEEP_EVALS = [0, 1, 9, 49, 99, 199, 299, 399, 496, 497]


def main():
    cfg = O.config_from_name("LightningDiT-S/1", input_size=16, num_classes=31)
    sd = O.init_state_dict(cfg, seed=0)
    n, K, scale, shift, P = 4, 31, 10.0, 0.3, 250
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(1234))
    y = torch.randint(0, K, (n,), generator=torch.Generator().manual_seed(4321))
    zz, yy = torch.cat([z, z], 0), torch.cat([y, torch.full((n,), K)])
    grid = O.time_grid(P, shift)
    t0 = time.time()
    final, trace = O.sample_ode(cfg, sd, zz, yy, grid, "heun", scale, return_trace=True)
    print(f"oracle: {len(trace)} evaluations in {time.time() - t0:.1f} s")
    out = dict(model="LightningDiT-S/1", n=n, num_classes=K, cfg_scale=scale, timestep_shift=shift, points=P,
               method="heun", seed_weights=0, seed_noise=1234, seed_labels=4321, grid=grid,
               final_kept=final[:n].clone(), keep_evals=KEEP_EVALS,
               velocities_kept=torch.stack([trace[i][:n] for i in KEEP_EVALS]).clone())
    path = os.path.join(ROOT, "tests", "golden", "long_s1_heun250.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
