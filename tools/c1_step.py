"""A few C1 training steps (256 x 200, H 256, bf16) for profiling:  ncu --metrics gpu__time_duration.sum ... python tools/c1_step.py [steps] [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, types
from adt_b200 import synth
from adt_b200.model import SASRecADT
from adt_b200.trainer import FusedTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else "C1"
cfg = synth.CONFIGS[name]
torch.manual_seed(0)
args = types.SimpleNamespace(device="cuda", num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"], dropout=cfg["p"])
m = SASRecADT(1, cfg["items"], args).cuda().train()
tr = FusedTrainer(m, [0.0124, 0.122], [0.0001, 0.05], weight_decay=cfg["wd"], seed=5, precision="bf16", use_graph=False)
rng = np.random.default_rng(3)
batch = synth.make_batch(rng, cfg)
for k in range(steps):
    tr.step(*batch)
torch.cuda.synchronize()
print("loss", tr.loss())
