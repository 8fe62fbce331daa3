"""Diagnostic: approx_kl / gate trajectory of consecutive update_params calls on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sard_b200 import synthetic
from sard_b200.config import DERL_YML, Config
from sard_b200.transform2act_agent import Transform2ActAgent
shape = synthetic.ENV_SHAPES['locomotion_ft']
cfg = Config(dict(DERL_YML))
torch.manual_seed(0); np.random.seed(1)
ag = Transform2ActAgent(cfg, torch.float32, 'cuda', 0, 1, env=synthetic.SyntheticEnv(shape))
roll = synthetic.generate_rollout(shape, 50000, seed=1234, dtype=np.float32)
for it in range(4):
    secs, log = ag.update_params(roll)
    pl = ag.last_update_stats['policy_log']
    print(f'update {it}: {secs:.3f}s valid {int(pl[:,3].sum())}/{len(pl)}  log={ {k: round(v,5) for k,v in log.items() if k in ("surr_loss","approx_kl","entropy")} }')
    print('  kl[0:12]   ', np.round(pl[:12, 2], 4))
    print('  kl[24:30]  ', np.round(pl[24:30, 2], 4), ' kl[-3:]', np.round(pl[-3:, 2], 4))
    print('  gradnorm[0:6]', np.round(pl[:6, 4], 4), 'vloss', np.round(ag.last_update_stats['value_loss'][[0, 1, 100, 239]], 4))
