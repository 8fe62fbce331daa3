"""Where does the end-to-end update spend its host-side time? (pack / upload / prepare / update, synchronised between stages)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic, packed  # noqa: E402
from sard_b200.config import DERL_YML, Config  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402

dev = torch.device('cuda', 0)
shape = synthetic.ENV_SHAPES[os.environ.get('ENV', 'locomotion_ft')]
d = dict(DERL_YML); d.update(min_batch_size=50000, mini_batch_size=2048, num_optim_epoch=10)
torch.manual_seed(0); np.random.seed(1234)
agent = Transform2ActAgent(Config(d), torch.float32, dev, 0, 1, env=synthetic.SyntheticEnv(shape))
roll = synthetic.generate_rollout(shape, 50000, seed=1234, dtype=np.float32)
batch = roll.to_traj_batch()
for _ in range(2):
    agent.update_params(batch)
torch.cuda.synchronize()
for it in range(3):
    tc0 = time.perf_counter()
    hnd = packed._hostpack.collect(batch.states, batch.actions)
    tc1 = time.perf_counter()
    del hnd
    print(f'  collect alone {1e3*(tc1-tc0):.1f} ms')
    t0 = time.perf_counter()
    ht = packed.host_tensors_from_lists(batch.states, batch.actions, batch.rewards, batch.masks, pin=True)
    t1 = time.perf_counter()
    arena = packed.PackedRollout(ht, dev, pin=True)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    agent._prepare(arena)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    agent.update_arena(arena)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f'pack(host) {1e3*(t1-t0):.1f} ms | arena ctor (enqueue) {1e3*(t2-t1):.1f} | upload drain {1e3*(t3-t2):.1f} | prepare {1e3*(t4-t3):.1f} | update_arena {1e3*(t5-t4):.1f} | total {1e3*(t5-t0):.1f}')
