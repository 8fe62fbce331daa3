"""The epoch driver of the drop-in agent (reference design_opt/agents/transform2act_agent.py:216-272, :411-455):
``agent.optimize(epoch)`` must run sample -> update_params -> evaluation -> subgroup-neighbour evaluation -> symmetry
update -> logging.  The CUDA path (``update_params``) is stubbed here; the GPU suite runs it for real."""
import numpy as np
import pytest
import torch

from sard_b200 import sampling, synthetic
from sard_b200.config import derl_config
from sard_b200.transform2act_agent import Transform2ActAgent


class StubPolicyForServer:
    """select_action with the policy's contract ([sum N, 8] float64 arrays) and a recording of the subgroup in force."""

    def __init__(self, group):
        self.group = group
        self.calls = []

    def select_action(self, states, mean_action=False):
        n = sum(int(np.asarray(s[3]).reshape(-1)[0]) for s in states)
        self.calls.append((self.group.cur_subgroup_name, bool(mean_action), len(states)))
        a = np.zeros((n, 8))
        a[:, 0] = 0.1
        return a, a.copy()


class Recorder:
    def __init__(self):
        self.lines, self.scalars = [], {}

    def info(self, s):
        self.lines.append(s)

    def add_scalar(self, k, v, x):
        self.scalars[k] = (float(v), x)


def make_agent():
    shape = synthetic.ENV_SHAPES['point_nav']
    cfg = derl_config(min_batch_size=120, eval_batch_size=60, mini_batch_size=60, num_optim_epoch=1)
    torch.manual_seed(0)
    ag = Transform2ActAgent(cfg, torch.float32, 'cpu', 0, 2, env=synthetic.SyntheticRolloutEnv(shape, exec_steps=6))
    return ag, shape


def test_optimize_runs_the_reference_epoch_sequence():
    ag, shape = make_agent()
    stub = StubPolicyForServer(ag.policy_net.group)
    ag.sampler = sampling.RolloutSampler(stub, lambda pid: synthetic.SyntheticRolloutEnv(shape, exec_steps=6, seed=pid), num_threads=2)
    ag.subgroup_eval_batch_size = 24
    seen = {}

    def fake_update(batch):                  # the CUDA path; checked against the oracle in the GPU suite
        seen['T'] = len(batch.states)
        seen['stages'] = sorted({int(s[2][0]) for s in batch.states})
        assert batch.masks.shape == batch.rewards.shape == (len(batch.states),)
        return 0.0, {'surr_loss': 0.0}
    ag.update_params = fake_update
    rec = Recorder()
    ag.logger, ag.tb_logger = rec, rec
    start = ag.policy_net.group.cur_subgroup_name
    neighbours = set(ag.policy_net.group.get_neibor_subgroup_name())
    np.random.seed(0)
    ag.optimize(0)
    assert seen['T'] >= 120 and seen['stages'] == [0, 1, 2]
    assert ag.total_time_steps >= 120
    # every neighbour subgroup was evaluated with ITS name in force, and never mixed with another in one forward
    served = {c[0] for c in stub.calls}
    assert neighbours <= served and start in served
    assert all(abs(ag.policy_net.group.value_subgroup[n]) > 0 for n in neighbours)
    # logging (agent :411-455)
    assert len(rec.lines) == 1 and 'T_update' in rec.lines[0]
    for k in ('train_R_avg', 'exec_R_eps_avg', 'train_surr_loss', 'train_random_prob', 'train_done_max_nsteps'):
        assert k in rec.scalars, k
    assert ag.save_best_flag and ag.best_rewards > -1000
    # the epsilon-greedy lattice step moved (or kept) the subgroup among the legal names
    assert ag.policy_net.group.cur_subgroup_name in neighbours | {start}


def test_sample_without_a_simulator_raises():
    shape = synthetic.ENV_SHAPES['point_nav']
    ag = Transform2ActAgent(derl_config(), torch.float32, 'cpu', 0, 1, env=synthetic.SyntheticEnv(shape))
    with pytest.raises(NotImplementedError):
        ag.sample(10)


def test_logger_and_batch_surface():
    env = synthetic.SyntheticRolloutEnv(synthetic.ENV_SHAPES['locomotion_ft'], exec_steps=4)
    stub = StubPolicyForServer(type('G', (), {'cur_subgroup_name': 'x'})())
    mem, lg = sampling.sample_worker(env, stub.select_action, 25, mean_action=True)
    merged = sampling.LoggerRLV1.merge([lg])
    batch = sampling.TrajBatchDisc([mem])
    assert merged.num_steps == len(batch.states) >= 25 and merged.num_episodes == merged.num_steps // 10
    assert merged.avg_exec_episode_reward == pytest.approx(4 * (1 - 0.01))
    assert batch.exps.sum() == 0 and batch.masks.sum() == merged.num_steps - merged.num_episodes
