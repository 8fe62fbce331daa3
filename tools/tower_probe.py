"""Times one critic update step by step on the GPU and prints block 0's phase stamps of the fused tower backward."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sard_b200 import _lib as L, synthetic
from sard_b200.config import Config, DERL_YML
from sard_b200.transform2act_agent import Transform2ActAgent

torch.cuda.set_device(0)
shape = synthetic.ENV_SHAPES['locomotion_ft']
T, M = 8192, 2048
roll = synthetic.generate_rollout(shape, T, seed=3, dtype=np.float32)
d = dict(DERL_YML); d.update(min_batch_size=T, mini_batch_size=M, num_optim_epoch=2)
ag = Transform2ActAgent(Config(d), torch.float32, 'cuda:0', 0, 1, env=synthetic.SyntheticEnv(shape))
for _ in range(2):
    secs, log = ag.update_params(roll)
torch.cuda.synchronize()
print('update_params', secs)
L.lib.sard_tt_debug_read.restype = C.c_int
L.lib.sard_tt_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 96)()
L.lib.sard_tt_debug_read(buf, 96)
t = list(buf)
names = {0: 'start', 1: 'row_setup done', 30: 'end'}
for i in range(3):
    names.update({2 + 6 * i: f'it{i} top', 3 + 6 * i: f'it{i} gather+put+arrive done', 4 + 6 * i: f'it{i} pair stores done',
                  5 + 6 * i: f'it{i} mma done', 6 + 6 * i: f'it{i} staged dagg + sync', 7 + 6 * i: f'it{i} gather dh done'})
for k in sorted(names):
    print(f'{names[k]:36s} +{t[k] - t[0]}')

print('--- forward (block 0)')
fn = {32: 'start', 33: 'in_fc operand + XH stores done', 34: 'in_fc mma done', 50: 'end'}
for l in range(3):
    fn.update({35 + 5 * l: f'L{l+1} top (prev stores done)', 36 + 5 * l: f'L{l+1} mma done', 37 + 5 * l: f'L{l+1} staged + sync',
               38 + 5 * l: f'L{l+1} gather+act done', 39 + 5 * l: f'L{l+1} put + sync2 done'})
for k in sorted(fn):
    print(f'{fn[k]:36s} +{t[k] - t[32]}')
print('--- wgrad (block 0)')
wn = {64: 'setup done', 65: 'pdl wait done', 66: 'epilogue warps waiting', 67: 'all MMAs done', 68: 'partials written'}
for k in sorted(wn):
    print(f'{wn[k]:36s} +{t[k] - t[64]}')
