"""torchrun check: Fit_GPcounts under G ranks returns, on every rank, the same table as a single-GPU run."""
import os, sys
import numpy as np, pandas as pd, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_batch
from gpcounts_b200 import Fit_GPcounts

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, Y = synth_batch(37, 200, seed=3)
cells = ["c%d" % i for i in range(200)]
Xdf = pd.DataFrame(X, index=cells); Ydf = pd.DataFrame(Y, index=["g%d" % i for i in range(37)], columns=cells)
df = Fit_GPcounts(Xdf, Ydf, sparse=True, M=16).One_sample_test("Negative_binomial")      # sharded: genes rank::world
dist.barrier()
dist.destroy_process_group()
ref = Fit_GPcounts(Xdf, Ydf, sparse=True, M=16).One_sample_test("Negative_binomial")     # unsharded on this GPU
same = np.array_equal(df.values, ref.values, equal_nan=True)
print("rank %d/%d sharded == single: %s (%d genes, bit for bit)" % (rank, world, same, len(df)))
sys.exit(0 if same else 1)
