"""BASELINE's target size on 8 GPUs in isolation (developer tool; torchrun): L=4, 4096 leaves, one warm-up + `steps`
timed build+solve steps, error against the analytic solution.  usage: torchrun ... tools/run_L4.py [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench  # noqa: E402
import jaxhps_b200 as hps  # noqa: E402
from jaxhps_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
_lib.load()
R = bench.Runner(torch, hps, 4, rank, world, dev, dist)
u = R.step(R.pb_res, R.g_dev, False)
err = R.error_vs_analytic(u)
del u
ms, wall, _ = R.timed(R.pb_res, R.g_dev, False, steps)
if rank == 0:
    print(json.dumps({"L": 4, "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "max_rel_error_vs_analytic_solution": err,
                      "ride_along": os.environ.get("HPS_DIST_RIDE", "auto")}), flush=True)
dist.barrier()
dist.destroy_process_group()
