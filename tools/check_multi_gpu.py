#!/usr/bin/env python
"""the sharded N-GPU result equals the single-GPU result on the same (z, y).
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multi_gpu.py
Rank 0 additionally samples the whole batch alone and compares with the all-gathered result (bit-exact: same
kernels, same per-sample arithmetic, chunk-invariant engine)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_COMPILE_DISABLE", "1")

import torch
import torch.distributed as dist

import va_vae_b200 as V
from va_vae_b200 import dist as vdist
from oracle import lightningdit_oracle as O   # weight recipe only


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name, K, n = "LightningDiT-S/1", 31, 16 * world + 3          # uneven shards on purpose
    cfg = O.config_from_name(name, input_size=16, num_classes=K)
    sd = O.init_state_dict(cfg, seed=0)
    m = V.LightningDiT_models[name](input_size=16, num_classes=K, in_channels=32, use_swiglu=True, use_rope=True,
                                    use_rmsnorm=True)
    m.load_state_dict(sd)
    m = m.eval().to(dev)
    eng = m.engine()
    grid = O.time_grid(8, 0.1)
    z = torch.randn(n, 32, 16, 16, generator=torch.Generator().manual_seed(1234)).to(dev)
    y = torch.randint(0, K, (n,), generator=torch.Generator().manual_seed(4321)).to(dev)

    def sample_local(zs, ys):
        zz, yy = vdist.build_cfg_batch(zs, ys, K)
        return eng.sample_ode(zz, yy, grid, "heun", True, 10.0, True, 0.11)[: zs.shape[0]]

    gathered = vdist.sample_sharded(sample_local, z, y)
    ok = True
    if rank == 0:
        alone = sample_local(z, y)
        diff = float((gathered - alone).abs().max())
        ok = torch.equal(gathered, alone)
        print(json.dumps(dict(world=world, n=n, bit_exact=ok, max_abs_diff=diff)), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
