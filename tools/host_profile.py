"""cProfile of the host side of one resident PPO update (where do the ~100 ms of issue time go?)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402

dev = torch.device('cuda', 0)
shape = synthetic.ENV_SHAPES['locomotion_ft']
d = dict(DERL_YML)
d.update(min_batch_size=50000, mini_batch_size=2048, num_optim_epoch=10)
torch.manual_seed(0); np.random.seed(1234)
agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape))
agent.two_thread_issue = False
roll = synthetic.generate_rollout(shape, 50000, seed=1234, dtype=np.float32)
arena = agent.pack(roll)
agent._prepare(arena)
for _ in range(2):
    agent.update_arena(arena)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
agent.update_arena(arena)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
