"""Host side of an epoch: rollout collection feeding ``update_params`` (reference
``khrylib/rl/agents/agent.py:109-135`` ``Agent.sample`` + ``design_opt/agents/transform2act_agent.py:56-109``
``sample_worker``), with the policy staying on the GPU behind :class:`~sard_b200.action_server.ActionServer`.

The simulator is external (MuJoCo on host threads) and not part of the CUDA path; what lives here is the glue the
drop-in agent needs so that ``agent.optimize(epoch)`` runs end to end:

* ``Memory`` / ``TrajBatchDisc`` / ``StatsLogger`` / ``LoggerRLV1`` with the reference's surface
  (``khrylib/utils/memory.py``, ``design_opt/utils/tools.py:4-16``, ``khrylib/utils/stats_logger.py``,
  ``design_opt/utils/logger.py:6-37``);
* ``RolloutSampler``: worker threads (or forked processes) stepping one env each; every ``select_action`` goes through
  the batching server, so the 20 workers of ``derl.yml`` share one batched GPU forward per window instead of 20 batch-1
  CPU forwards;
* ``sample_many``: several subgroup variants sampled CONCURRENTLY through the same server (section 8f.3: the reference
  walks the 6+ neighbour subgroups one after the other, ``transform2act_agent.py:244-256``).  Requests carry the
  subgroup name; the server never mixes two subgroups (or two sampling modes) in one forward.
"""
from __future__ import annotations

import math
import threading
import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


class StatsLogger:
    """khrylib/utils/stats_logger.py."""

    def __init__(self, is_nparray=False):
        self.is_nparray = is_nparray
        self.total_val = 0
        self.min_val = math.inf
        self.max_val = -math.inf
        self.n = 0

    def log(self, val):
        self.total_val += val
        self.min_val = np.minimum(self.min_val, val) if self.is_nparray else min(self.min_val, val)
        self.max_val = np.maximum(self.max_val, val) if self.is_nparray else max(self.max_val, val)
        self.n += 1

    def avg(self):
        return 0 if self.n == 0 else self.total_val / self.n

    def total(self):
        return self.total_val

    def min(self):
        return self.min_val

    def max(self):
        return self.max_val

    @classmethod
    def merge(cls, loggers):
        out = cls(is_nparray=loggers[0].is_nparray)
        out.total_val = sum(x.total_val for x in loggers)
        out.min_val = np.min(np.stack([x.min_val for x in loggers]), axis=0)
        out.max_val = np.max(np.stack([x.max_val for x in loggers]), axis=0)
        out.n = sum(x.n for x in loggers)
        return out


class LoggerRLV1:
    """Reward bookkeeping of one sampler (khrylib/rl/core/logger_rl.py:5-77 + design_opt/utils/logger.py:6-37)."""

    DONE = ('done_isfinite', 'done_min_height', 'done_max_height', 'done_angle', 'done_max_nsteps')

    def __init__(self, init_stats_logger=True, use_c_reward=False):
        self.use_c_reward = use_c_reward
        self.num_steps = 0
        self.num_episodes = 0
        self.sample_time = 0
        self.stats_names = ['episode_len', 'reward', 'episode_reward', *self.DONE, 'exec_reward', 'exec_episode_reward']
        if init_stats_logger:
            self.stats_loggers = {x: StatsLogger() for x in self.stats_names}

    def start_episode(self, env):
        self.episode_len = 0
        self.episode_reward = 0
        self.exec_episode_reward = 0

    def step(self, env, reward, c_reward, c_info, info):
        self.episode_len += 1
        self.episode_reward += reward
        self.stats_loggers['reward'].log(reward)
        if info.get('done', False):
            for k, v in info.items():
                if k in self.DONE:
                    self.stats_loggers[k].log(v)
        if info.get('stage') == 'execution':
            self.stats_loggers['exec_reward'].log(reward)
            self.exec_episode_reward += reward

    def end_episode(self, env):
        self.num_steps += self.episode_len
        self.num_episodes += 1
        self.stats_loggers['episode_len'].log(self.episode_len)
        self.stats_loggers['episode_reward'].log(self.episode_reward)
        self.stats_loggers['exec_episode_reward'].log(self.exec_episode_reward)

    def end_sampling(self):
        pass

    @classmethod
    def merge(cls, logger_list, **kwargs):
        lg = cls(init_stats_logger=False, **kwargs)
        lg.num_episodes = sum(x.num_episodes for x in logger_list)
        lg.num_steps = sum(x.num_steps for x in logger_list)
        lg.stats_loggers = {s: StatsLogger.merge([x.stats_loggers[s] for x in logger_list]) for s in lg.stats_names}
        sl = lg.stats_loggers
        lg.total_reward = sl['reward'].total()
        lg.avg_episode_len = sl['episode_len'].avg()
        lg.avg_episode_reward = sl['episode_reward'].avg()
        lg.max_episode_reward = sl['episode_reward'].max()
        lg.min_episode_reward = sl['episode_reward'].min()
        lg.max_reward, lg.min_reward, lg.avg_reward = sl['reward'].max(), sl['reward'].min(), sl['reward'].avg()
        lg.avg_exec_reward = sl['exec_reward'].avg()
        lg.avg_exec_episode_reward = sl['exec_episode_reward'].avg()
        return lg


class Memory:
    """khrylib/utils/memory.py (the parts the sampler uses)."""

    def __init__(self):
        self.memory = []

    def push(self, *args):
        self.memory.append([*args])

    def sample(self, batch_size=None):
        return self.memory

    def append(self, other):
        self.memory += other.memory

    def __len__(self):
        return len(self.memory)


class TrajBatchDisc:
    """design_opt/utils/tools.py:4-16: lists of per-step states / actions + stacked masks, rewards, exps."""

    def __init__(self, memory_list):
        memory = memory_list[0]
        for x in memory_list[1:]:
            memory.append(x)
        cols = list(zip(*memory.sample()))
        self.states = list(cols[0])
        self.actions = list(cols[1])
        self.masks = np.stack(cols[2])
        self.next_states = list(cols[3])
        self.rewards = np.stack(cols[4])
        self.exps = np.stack(cols[5])


def sample_worker(env, select_action: Callable, min_batch_size: int, mean_action: bool, noise_rate: float = 1.0,
                  reset_args: Optional[dict] = None, rng: Optional[np.random.Generator] = None, max_episode_steps: int = 10000):
    """One sampler (agent :56-100): episodes until ``min_batch_size`` steps; returns ``(Memory, LoggerRLV1)``.
    ``select_action([state], mean_action) -> (action, action_env)`` is the policy call (the server's client)."""
    memory, logger = Memory(), LoggerRLV1()
    rng = rng if rng is not None else np.random.default_rng()
    while logger.num_steps < min_batch_size:
        state = env.reset(extra_args=reset_args) if reset_args is not None else env.reset()
        logger.start_episode(env)
        for _ in range(max_episode_steps):
            use_mean = bool(mean_action) or bool(rng.random() < 1.0 - noise_rate)
            action, action_env = select_action([state], use_mean)
            next_state, reward, done, info = env.step(action_env)
            logger.step(env, reward, 0.0, np.array([0.0]), info)
            memory.push(state, action, 0 if done else 1, next_state, reward, 1 - int(use_mean))
            if done:
                break
            state = next_state
        logger.end_episode(env)
    logger.end_sampling()
    return memory, logger


class RolloutSampler:
    """``sample()`` for the drop-in agent: ``nthreads`` workers, one env each, one batching action server.

    ``env_factory(pid) -> env`` builds a worker's env (the reference forks the agent and with it ``self.env``; threads
    need one instance each).  ``policy`` needs ``select_action(states, mean_action)`` and, for subgroup variants,
    ``group.cur_subgroup_name`` / ``set_cur_subgroup_name`` / ``get_cur_num``.
    """

    def __init__(self, policy, env_factory: Callable[[int], object], num_threads: int = 1, max_batch: int = 64,
                 max_wait_ms: float = 1.0, noise_rate: float = 1.0, seed: int = 0):
        self.policy = policy
        self.env_factory = env_factory
        self.num_threads = max(1, int(num_threads))
        self.max_batch, self.max_wait_ms = max_batch, max_wait_ms
        self.noise_rate = noise_rate
        self.seed = seed
        self._envs: Dict[int, object] = {}
        self.last_server_stats = {}

    def _env(self, slot: int):
        if slot not in self._envs:
            self._envs[slot] = self.env_factory(slot)
        return self._envs[slot]

    def sample(self, min_batch_size: int, mean_action: bool = False, nthreads: Optional[int] = None):
        out = self.sample_many({None: min_batch_size}, mean_action, nthreads)
        return out[None]

    def sample_many(self, jobs: Dict[Optional[str], int], mean_action: bool = False, nthreads: Optional[int] = None):
        """``jobs``: subgroup name (None = the current one) -> min_batch_size.  All jobs run concurrently, ``nthreads``
        workers each, through one server; returns name -> ``(TrajBatchDisc, LoggerRLV1)``."""
        from .action_server import ActionServer
        nthreads = self.num_threads if nthreads is None else max(1, int(nthreads))
        group = getattr(self.policy, 'group', None)
        server = ActionServer(self.policy, max_batch=self.max_batch, max_wait_ms=self.max_wait_ms)
        t0 = time.time()
        results: Dict[Optional[str], List] = {name: [None] * nthreads for name in jobs}
        errors: List[BaseException] = []
        threads = []
        slot = 0
        for name, min_batch in jobs.items():
            per_thread = int(math.floor(min_batch / nthreads))
            reset_args = None
            if group is not None:
                prev = group.cur_subgroup_name
                if name is not None:
                    group.set_cur_subgroup_name(name)
                reset_args = {'num': group.get_cur_num()}
                group.set_cur_subgroup_name(prev)
            for i in range(nthreads):
                client = server.client(key=name)
                env = self._env(slot)
                rng = np.random.default_rng([self.seed, slot, int(t0 * 1e3) & 0xffff])

                def work(name=name, i=i, env=env, client=client, rng=rng, reset_args=reset_args, per_thread=per_thread):
                    try:
                        results[name][i] = sample_worker(env, client.select_action, per_thread, mean_action, self.noise_rate,
                                                         reset_args, rng)
                    except BaseException as e:          # noqa: BLE001
                        errors.append(e)
                th = threading.Thread(target=work, daemon=True)
                threads.append(th)
                slot += 1
        server.start()
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        server.stop()
        self.last_server_stats = {'batches': server.batches, 'served': server.served,
                                  'mean_batch': server.served / max(server.batches, 1)}
        if errors:
            raise errors[0]
        out = {}
        for name, res in results.items():
            batch = TrajBatchDisc([m for m, _ in res])
            logger = LoggerRLV1.merge([lg for _, lg in res])
            logger.sample_time = time.time() - t0
            out[name] = (batch, logger)
        return out
