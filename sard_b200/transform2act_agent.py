"""Drop-in ``Transform2ActAgent`` PPO driver (reference: design_opt/agents/transform2act_agent.py
on top of khrylib/rl/agents/{agent,agent_pg,agent_ppo}.py).

The learning half (``update_params`` -> value pre-pass -> GAE -> ``update_policy`` with its
minibatch loop of ``update_value`` + ``ppo_loss`` + clip + Adam) runs on the device through
:mod:`sard_b200.engine`; rollouts (``sample``), logging and checkpoint I/O are host code with the
reference's behaviour.  The KL/loss gate is evaluated on the device, so the minibatch loop has no
host synchronisation; the per-step logs are read back once at the end.
"""
from __future__ import annotations

import math
import os
import pickle
import time
from collections import Counter
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .engine import Comm
from .nets import AdamState
from .packed import EpochPlan, Ordering, PackedRollout, host_tensors_from_lists, stage_partition
from .rl_core import estimate_advantages
from .synthetic import HostRollout
from .transform2act_critic import Transform2ActValue
from .transform2act_policy import Transform2ActPolicy

EVAL_CHUNK = 16384   # graphs per eval-mode launch (the reference uses 10000; results do not depend on it)


def tensorfy(np_list, device=torch.device('cpu')):
    """Reference helper kept for API compatibility (agent :20-24)."""
    if isinstance(np_list[0], list):
        return [[torch.tensor(x).to(device) if i in {0, 1, 5, 6, 7} else x for i, x in enumerate(y)] for y in np_list]
    return [torch.tensor(y).to(device) for y in np_list]


class Transform2ActAgent:
    """``Transform2ActAgent(cfg, dtype, device, seed, num_threads, training=True, checkpoint=0)``.

    ``env`` may be passed explicitly (any object with the attributes read in ``setup_env``); when
    omitted the reference's ``design_opt.envs.env_dict`` is looked up, which needs the reference's
    MuJoCo stack on the host.
    """

    def __init__(self, cfg, dtype, device, seed, num_threads, training=True, checkpoint=0, env=None, comm=None):
        self.cfg = cfg
        self.training = training
        self.device = torch.device(device)
        self.dtype = dtype
        self.loss_iter = 0
        self.comm: Optional[Comm] = comm
        self.setup_env(env)
        self.seed(seed)
        self.setup_logger()
        self.setup_policy()
        self.setup_value()
        self.setup_optimizer()
        self.scheduled_params = {}
        self.gamma, self.tau, self.clip_epsilon = cfg.gamma, cfg.tau, cfg.clip_epsilon
        self.opt_num_epochs = cfg.num_optim_epoch
        self.mini_batch_size = cfg.mini_batch_size
        self.use_mini_batch = cfg.mini_batch_size < cfg.min_batch_size
        self.policy_grad_clip = [(self.policy_net.parameters(), 40)]
        # The reference's clip list holds a generator: only its first use clips (agent_ppo.py:53-56).
        self.clip_first_step_only = True
        # zero_grad() semantics of the host torch the run is meant to reproduce: 'none' (torch >= 2: gradients of a stage
        # absent from a minibatch are None and Adam skips its parameters) or 'zeros' (torch < 2.0, the reference's pinned
        # 1.8: they are zero tensors and Adam keeps stepping those parameters from momentum); see INTEGRATION.md
        self.zero_grad_semantics = os.environ.get('SARD_ZERO_GRAD', 'none')
        self.value_opt_niter = 1
        self.num_threads = num_threads
        self.noise_rate = 1.0
        self.running_state = None
        self.custom_reward = None
        self.end_reward = False
        self.sample_modules = [self.policy_net]
        self.update_modules = [self.policy_net, self.value_net]
        self.enable_infer = cfg.enable_infer
        self.total_time_steps = 0
        self.last_update_stats = {}
        self.sampler = None                 # rollout source: RolloutSampler-like object with sample() / sample_many()
        self.subgroup_eval_batch_size = 1000        # agent :252
        self.prefetch_gather = os.environ.get('SARD_PREFETCH', '1') != '0'
        self._gather_ctx = None
        self.overlap_nets = True          # critic and policy updates of a minibatch on two CUDA streams
        # ... optionally issued from two host threads (the C++ schedules release the GIL).  Measured SLOWER on one B200
        # box (119 -> 137 ms/update: the GIL hand-offs cost more than the parallel issue gains), so it is off by default
        self.two_thread_issue = os.environ.get('SARD_TWO_THREADS', '0') != '0'
        self._streams = None
        if checkpoint != 0:
            self.load_checkpoint(checkpoint)

    # ---- setup ------------------------------------------------------------------------------
    def setup_env(self, env=None):
        if env is None:
            try:
                from design_opt.envs import env_dict          # the reference's MuJoCo envs, if installed
            except Exception as e:                             # noqa: BLE001
                raise RuntimeError('no env given and design_opt.envs is not importable; rollouts stay on the '
                                   "reference's host stack (see INTEGRATION.md)") from e
            env = env_dict[self.cfg.env_name](self.cfg, self)
        self.env = env
        self.attr_fixed_dim = env.attr_fixed_dim
        self.attr_design_dim = env.attr_design_dim
        self.sim_obs_dim = env.sim_obs_dim
        self.state_dim = self.attr_fixed_dim + self.sim_obs_dim + self.attr_design_dim
        self.control_action_dim = env.control_action_dim
        self.skel_num_action = env.skel_num_action
        self.action_dim = self.control_action_dim + self.attr_design_dim
        self.sim_obs_hfield_dim = env.sim_obs_hfield_dim
        self.sim_obs_obj_dim = env.sim_obs_obj_dim
        self.sim_obs_goal_dim = env.sim_obs_goal_dim

    def seed(self, seed):
        if hasattr(self.env, 'seed'):
            self.env.seed(seed)

    def setup_logger(self):
        self.tb_logger = None
        self.logger = None
        self.best_rewards = -1000.0
        self.save_best_flag = False

    def setup_policy(self):
        self.policy_net = Transform2ActPolicy(self.cfg.policy_specs, self)
        self.policy_net.to(self.device)

    def setup_value(self):
        self.value_net = Transform2ActValue(self.cfg.value_specs, self)
        self.value_net.to(self.device)

    def setup_optimizer(self):
        cfg = self.cfg
        if cfg.policy_optimizer != 'Adam' or cfg.value_optimizer != 'Adam':
            raise NotImplementedError('only Adam is on the CUDA path (design_opt/cfg/derl.yml)')
        if cfg.policy_weightdecay != 0.0 or cfg.value_weightdecay != 0.0:
            raise NotImplementedError('weight decay would touch never-seen IndexLinear rows; not on the CUDA path')
        self.optimizer_policy = AdamState(cfg.policy_lr)
        self.optimizer_value = AdamState(cfg.value_lr)

    # ---- checkpoint I/O (agent :167-210) ---------------------------------------------------------
    def save_checkpoint_to(self, cp_path, epoch=0):
        cp = {'policy_dict': {k: v.detach().cpu() for k, v in self.policy_net.state_dict().items()},
              'value_dict': {k: v.detach().cpu() for k, v in self.value_net.state_dict().items()},
              'running_state': self.running_state, 'loss_iter': self.loss_iter, 'best_rewards': self.best_rewards,
              'epoch': epoch, 'symmetry': self.policy_net.group.cur_subgroup_name if self.enable_infer else None}
        # beyond the reference's keys (agent :192-199 drops the optimiser): Adam moments and step counters, the
        # IndexLinear tables' moments stored sparsely (live rows only); ignored by the reference's load_checkpoint
        cp['optimizer_state'] = {
            'policy': self.optimizer_policy.state_dict(self.policy_net, self.policy_net.engine.store),
            'value': self.optimizer_value.state_dict(self.value_net, self.value_net.engine.store)}
        with open(cp_path, 'wb') as f:
            pickle.dump(cp, f)

    def save_checkpoint(self, epoch):
        cfg = self.cfg
        extra = cfg.agent_specs.get('additional_saves', None)
        if (cfg.save_model_interval > 0 and (epoch + 1) % cfg.save_model_interval == 0) or \
           (extra is not None and (epoch + 1) % extra[0] == 0 and epoch + 1 <= extra[1]):
            self.save_checkpoint_to('%s/epoch_%04d.p' % (cfg.model_dir, epoch + 1), epoch)
        if self.save_best_flag:
            self.save_checkpoint_to('%s/best.p' % cfg.model_dir, epoch)

    def load_checkpoint(self, checkpoint):
        cfg = self.cfg
        if isinstance(checkpoint, int):
            cp_path = '%s/epoch_%04d.p' % (cfg.model_dir, checkpoint)
        elif os.path.exists(str(checkpoint)):
            cp_path = str(checkpoint)
        else:
            cp_path = '%s/%s.p' % (cfg.model_dir, checkpoint)
        with open(cp_path, 'rb') as f:
            cp = pickle.load(f)
        self.policy_net.load_state_dict(cp['policy_dict'])
        self.value_net.load_state_dict(cp['value_dict'])
        self.running_state = cp['running_state']
        self.loss_iter = cp['loss_iter']
        self.best_rewards = cp.get('best_rewards', self.best_rewards)
        ost = cp.get('optimizer_state')
        if ost is not None and self.policy_net.control_action_log_std.device.type == 'cuda':
            with torch.cuda.device(self.device):
                self.policy_net.prepare(None, self.optimizer_policy)         # build the flat stores the moments map onto
                self.value_net.prepare(self.optimizer_value)
                self.optimizer_policy.load_state_dict(ost['policy'], self.policy_net, self.policy_net.engine.store)
                self.optimizer_value.load_state_dict(ost['value'], self.value_net, self.value_net.engine.store)
        if self.enable_infer and cp.get('symmetry') is not None:
            self.policy_net.group.cur_subgroup_name = cp['symmetry']

    def pre_epoch_update(self, epoch):
        for p in self.scheduled_params.values():
            p.set_epoch(epoch)

    # ---- epoch driver (agent :216-272, :411-455) --------------------------------------------------------------
    def optimize(self, epoch):
        """agent :216-219."""
        self.pre_epoch_update(epoch)
        info = self.optimize_policy(epoch)
        self.log_optimize_policy(epoch, info)

    def optimize_policy(self, epoch):
        """agent :221-272: sample -> update_params -> evaluation rollouts -> subgroup-neighbour evaluation ->
        symmetry update.  The neighbour subgroups are sampled concurrently through one action server
        (``sample_subgroups``) instead of one after the other (agent :250-253)."""
        cfg = self.cfg
        t0 = time.time()
        batch, log = self.sample(cfg.min_batch_size)
        self.total_time_steps += log.num_steps
        t1 = time.time()
        _, log_train = self.update_params(batch)
        t2 = time.time()
        _, log_eval = self.sample(cfg.eval_batch_size, mean_action=True)
        t3 = time.time()
        if self.enable_infer:
            group = self.policy_net.group
            group.store_cur_subgroup_name()
            if getattr(cfg, 'enable_design', True) and cfg.updating_subgroup:
                if cfg.struct_subgroup:
                    evaluate_names = list(group.get_neibor_subgroup_name())
                else:
                    evaluate_names = list(np.random.choice(group.get_neibor_subgroup_name(), size=2, replace=False))
            else:
                evaluate_names = []
            value_dict = {name: lg.avg_episode_reward
                          for name, lg in self.sample_subgroups(evaluate_names, self.subgroup_eval_batch_size).items()}
            group.update_all_values(value_dict)
            group.restore_cur_subgroup_name()
            self.policy_net.update_symmetry(log.avg_episode_reward)
            for k, v in self.policy_net.get_info().items():
                if isinstance(v, dict):
                    for k1, v1 in v.items():
                        log_train[f'{k}_{k1}'] = v1
                else:
                    log_train[f'{k}'] = v
        return {'log': log, 'log_eval': log_eval, 'T_sample': t1 - t0, 'T_update': t2 - t1, 'T_eval': t3 - t2,
                'T_total': t3 - t0, 'log_train': log_train}

    def log_optimize_policy(self, epoch, info):
        """agent :411-455: console line, best-reward flag, TensorBoard / wandb scalars when a writer is attached."""
        cfg = self.cfg
        log, log_eval, log_train = info['log'], info['log_eval'], info['log_train']
        if self.logger is not None:
            self.logger.info(f'{epoch}\tT_sample {info["T_sample"]:.2f}\tT_update {info["T_update"]:.2f}\tT_eval {info["T_eval"]:.2f}\t'
                             f'train_R {log.avg_reward:.2f}\ttrain_R_eps {log.avg_episode_reward:.2f}\t'
                             f'exec_R {log_eval.avg_exec_reward:.2f}\texec_R_eps {log_eval.avg_exec_episode_reward:.2f}\t'
                             f'{getattr(cfg, "id", "")}')
        if log_eval.avg_exec_episode_reward > self.best_rewards:
            self.best_rewards = log_eval.avg_exec_episode_reward
            self.save_best_flag = True
        else:
            self.save_best_flag = False
        tb = self.tb_logger
        if tb is None:
            return
        tb_x = self.total_time_steps
        for name, v in (('train_R_avg', log.avg_reward), ('train_R_eps_avg', log.avg_episode_reward),
                        ('eval_R_eps_avg', log_eval.avg_episode_reward), ('exec_R_avg', log_eval.avg_exec_reward),
                        ('exec_R_eps_avg', log_eval.avg_exec_episode_reward)):
            tb.add_scalar(name + '_t', v, tb_x)
            tb.add_scalar(name, v, epoch)
        for k, v in log.stats_loggers.items():
            if 'done_' in k:
                tb.add_scalar(f'train_{k}', v.avg(), epoch)
        for k, v in log_train.items():
            tb.add_scalar(f'train_{k}', v, epoch)

    # ---- rollouts: the simulator is external host code; the policy stays on the GPU behind the action server ----
    def _sampler(self):
        if self.sampler is None:
            env = self.env
            if not (hasattr(env, 'reset') and hasattr(env, 'step')):
                raise NotImplementedError('this env has no reset()/step(): attach a rollout source with agent.sampler = '
                                          'RolloutSampler(...) or feed update_params() a TrajBatchDisc-like batch '
                                          '(see INTEGRATION.md)')
            from .sampling import RolloutSampler
            factory = getattr(self, 'env_factory', None)
            if factory is None:
                if self.num_threads > 1:
                    raise NotImplementedError('num_threads > 1 needs agent.env_factory(pid) -> env (one simulator per worker)')
                factory = lambda pid: env
            self.sampler = RolloutSampler(self.policy_net, factory, self.num_threads, noise_rate=self.noise_rate)
        return self.sampler

    def sample(self, min_batch_size, mean_action=False, render=False, nthreads=None):
        """agent.py:109-135: ``(TrajBatchDisc, LoggerRLV1)``.  Workers call ``select_action`` through the batching
        server (one batched GPU forward per window) instead of moving the nets to the CPU (agent.py:114-115)."""
        for m in self.sample_modules:
            m.train(False)
        return self._sampler().sample(min_batch_size, mean_action, nthreads)

    def sample_subgroups(self, names, min_batch_size, mean_action=False):
        """Rollouts under each of the subgroups ``names`` (agent :250-253), all at once: name -> LoggerRLV1."""
        if not names:
            return {}
        for m in self.sample_modules:
            m.train(False)
        out = self._sampler().sample_many({n: min_batch_size for n in names}, mean_action)
        return {n: lg for n, (_, lg) in out.items()}

    # ---- the hot path ------------------------------------------------------------------------------
    def pack(self, batch) -> PackedRollout:
        """``TrajBatchDisc`` (or a ``HostRollout``) -> device arena: one packed upload instead of tensorfy."""
        if isinstance(batch, HostRollout):
            return PackedRollout(batch, self.device, pin=True)
        return PackedRollout(host_tensors_from_lists(batch.states, batch.actions, batch.rewards, batch.masks, pin=True),
                             self.device, pin=True)

    def _prepare(self, arena: PackedRollout):
        self.policy_net.engine.comm = self.value_net.engine.comm = self.comm
        self.policy_net.prepare(arena, self.optimizer_policy, comm=self.comm)
        self.value_net.prepare(self.optimizer_value)
        if self.zero_grad_semantics not in ('none', 'zeros'):
            raise ValueError("zero_grad_semantics must be 'none' or 'zeros'")
        self.policy_net.engine.zero_grad_semantics = self.zero_grad_semantics

    def _eval_values(self, arena: PackedRollout):
        order = Ordering(arena, np.arange(arena.T))
        eng = self.value_net.engine
        for lo in range(0, arena.T, EVAL_CHUNK):
            eng.run(arena, order.run(lo, min(arena.T, lo + EVAL_CHUNK)), 0)
        return order

    def _eval_log_probs(self, arena: PackedRollout):
        order = Ordering(arena, stage_partition(arena.stage_np))
        eng = self.policy_net.engine
        bounds = np.searchsorted(order.stage_np, [0, 1, 2, 3])
        for s in (2, 1, 0):
            for lo in range(int(bounds[s]), int(bounds[s + 1]), EVAL_CHUNK):
                eng.run_stage(s, arena, order.run(lo, min(int(bounds[s + 1]), lo + EVAL_CHUNK)), 0)
        return order

    def update_params(self, batch):
        """agent :274-303.  Returns ``(seconds, log)`` like the reference."""
        t0 = time.time()
        for m in self.update_modules:
            m.train(True)
        with torch.cuda.device(self.device):
            arena = self.pack(batch)
            self._prepare(arena)
            log = self.update_arena(arena)
        return time.time() - t0, log

    def _net_streams(self):
        if self._streams is None or self._streams[0].device != self.device:
            # (critic, policy) launching streams; SARD_PRIO_VALUE / SARD_PRIO_POLICY: CUDA stream priorities (0 = default,
            # negative = higher); the policy's chain is the longer one (85 vs 70 ms alone), so its kernels go first
            pv, pp = int(os.environ.get('SARD_PRIO_VALUE', 0)), int(os.environ.get('SARD_PRIO_POLICY', -1))
            self._streams = (torch.cuda.Stream(self.device, priority=pv), torch.cuda.Stream(self.device, priority=pp))
        return self._streams

    def update_arena(self, arena: PackedRollout):
        """``update_params`` after packing, on a resident arena: value pre-pass, GAE, fixed log-prob pre-pass (agent
        :284-296, :316-327), then the minibatch loop.  The two pre-passes are independent (critic / policy, disjoint
        outputs), so the policy's runs on the policy stream beside the critic's and the GAE."""
        for m in self.update_modules:
            m.train(False)
        # the reference clips only the first policy step of a run (its clip list holds a generator, agent :48): once
        # that step is behind us the gradient norm is dead work.  One 4-byte read while the GPU is idle.
        opt = self.optimizer_policy
        armed = True if opt.ctl_i is None else bool(int(opt.ctl_i[1].item()))
        self.policy_net.engine.skip_sqnorm = bool(self.clip_first_step_only and not armed)
        fixed_ready = False
        if self.overlap_nets:
            main = torch.cuda.current_stream(self.device)
            sp = self._net_streams()[1]
            sp.wait_stream(main)
            with torch.cuda.stream(sp):
                self._fixed_order = self._eval_log_probs(arena)      # keep the ordering tensors alive
            fixed_ready = True
        self._eval_values(arena)
        for m in self.update_modules:
            m.train(True)
        adv, ret = estimate_advantages(arena.rewards, arena.masks, arena.values.unsqueeze(1), self.gamma, self.tau)
        arena.advantages.copy_(adv.reshape(-1))
        arena.returns.copy_(ret.reshape(-1))
        if self.cfg.agent_specs.get('reinforce', False):
            arena.advantages.copy_(arena.returns)
        return self.update_policy_packed(arena, fixed_ready=fixed_ready)

    def update_policy(self, states, actions, returns, advantages, exps):
        """Reference signature (agent :313): lists of states/actions plus ``[T,1]`` tensors."""
        with torch.cuda.device(self.device):
            arena = PackedRollout.from_lists(states, actions, device=self.device)
            self._prepare(arena)
            arena.returns.copy_(returns.reshape(-1).to(self.device, torch.float32))
            arena.advantages.copy_(advantages.reshape(-1).to(self.device, torch.float32))
            return self.update_policy_packed(arena)

    def draw_perm(self, T: int) -> np.ndarray:
        perm = np.arange(T)
        np.random.shuffle(perm)          # same RNG call as the reference (agent :333)
        return perm

    def update_policy_packed(self, arena: PackedRollout, perms: Optional[List[np.ndarray]] = None, fixed_ready: bool = False):
        comm = self.comm
        world = comm.world if comm is not None else 1
        peng, veng = self.policy_net.engine, self.value_net.engine
        t_host0 = time.perf_counter()
        if not fixed_ready:                  # fixed_ready: update_arena already queued the pre-pass on the policy stream
            for m in self.update_modules:
                m.train(False)
            self._eval_log_probs(arena)
            for m in self.update_modules:
                m.train(True)
        T = arena.T
        M = self.mini_batch_size if self.use_mini_batch else T
        if not self.use_mini_batch:
            # the reference's full-batch branch (agent :361-367) ignores valid_loss and always steps; the device gate of
            # this path would skip such a step
            raise NotImplementedError('mini_batch_size >= min_batch_size (the full-batch branch of update_policy) is not '
                                      'on the CUDA path: it steps regardless of the KL gate; derl.yml uses minibatches')
        batch_design = bool(self.cfg.agent_specs.get('batch_design', False))
        n_mb = T // M
        if world > 1:
            # rollouts end on episode boundaries, so T differs between ranks: every rank must run the same number of
            # minibatch steps (= collectives); the ranks with a longer rollout drop their surplus minibatches
            n_mb = comm.min_int(n_mb)
            if n_mb <= 0:
                raise ValueError('data-parallel update: a rank holds fewer transitions than one minibatch')
        n_steps = self.opt_num_epochs * n_mb
        plans_alive = []
        plog = torch.zeros(max(n_steps, 1), 8, dtype=torch.float32, device=self.device)
        vlog = torch.zeros(max(n_steps, 1), 8, dtype=torch.float32, device=self.device)
        step = 0
        h2d = 0
        cur_perm = np.arange(T)
        # The critic update and the policy update of a minibatch are independent (disjoint parameters,
        # workspaces and gradient buckets; the arena is read-only), and each one alone launches grids of
        # only ~150-300 CTAs on 148 SMs: run them on two streams so their kernels share the machine.
        main = torch.cuda.current_stream(self.device)
        sv, sp = self._net_streams()
        sv_h, sp_h = sv.cuda_stream, sp.cuda_stream
        lanes = comm.lanes(4) if (world > 1 and os.environ.get('SARD_COMM_LANES', '1') != '0') else [comm] * 4
        if self.overlap_nets and self.prefetch_gather:
            veng.GA.zero_(); peng.GA.zero_()         # the fused optimiser tail keeps the buckets zeroed from here on
        sv.wait_stream(main); sp.wait_stream(main)
        if world == 1 and self.overlap_nets and self.prefetch_gather and self.two_thread_issue and n_steps > 0:
            step, h2d, cur_perm = self._update_two_threads(arena, perms, T, M, batch_design, n_mb, vlog, plog, main, sv, sp)
            main.wait_stream(sv); main.wait_stream(sp)
            return self._finish_update(arena, plog, vlog, step, h2d, t_host0, cur_perm)
        for ep in range(self.opt_num_epochs):
            # like the reference, each epoch permutes the already permuted lists (agent :331-342)
            p = perms[ep] if perms is not None else self.draw_perm(T)
            cur_perm = cur_perm[p]
            plan = EpochPlan(arena, cur_perm, M, batch_design, n_mb=n_mb)
            cur_perm = plan.perm
            h2d += plan.h2d_bytes
            sv.wait_stream(main); sp.wait_stream(main)       # the ordering upload happened on the main stream
            counts = np.concatenate([plan.stage_graphs, plan.stage_nodes], axis=1)      # [n_mb, 6]
            if world > 1:
                # host metadata travels over the host-side group: the device round trip (pageable H2D + NCCL + D2H on the
                # main stream) cost 2 ms per epoch on 2 GPUs
                counts = comm.sum_array(counts)
            for k in range(plan.n_mb):
                B_glob = int(counts[k, :3].sum())
                active = [bool(counts[k, s] > 0) for s in range(3)]
                vrun, pruns = plan.value_run(k), plan.policy_runs(k)
                skel_nodes = int(counts[k, 3])
                if world > 1 and self.overlap_nets and self.prefetch_gather:
                    # as below, and the moments all-gathers of minibatch k+1 leave the critical path with the gather.
                    # Each of the four chains of collectives has its own communicator (Comm.lanes); every rank issues
                    # the same sequence on each of them.
                    slot = k & 1
                    act_of = lambda j: [bool(counts[j, s] > 0) for s in range(3)]
                    c_vr, c_pr, c_vg, c_pg = lanes
                    if k == 0:
                        self._issue_prefetch(arena, plan, 0, 0, main, (c_vg, c_pg), act_of(0))
                    if k + 1 < plan.n_mb:
                        self._issue_prefetch(arena, plan, k + 1, slot ^ 1, None, (c_vg, c_pg), act_of(k + 1))
                    gv, gp, ready, free = self._gather_ctx
                    sv.wait_event(ready[0][slot])
                    veng.compute_prefetched(arena, vrun, B_glob, slot, world, st=sv_h)
                    sp.wait_event(ready[1][slot])
                    peng.compute_prefetched(arena, pruns, B_glob, self.clip_epsilon, slot, world, active, st=sp_h)
                    if veng.p2p is not None:                 # one kernel over peer memory on the net's own stream
                        veng.dp_allreduce(c_vr, st=sv_h)
                    else:
                        with torch.cuda.stream(sv):
                            veng.dp_allreduce(c_vr)
                    veng.adam(B_glob, vlog[step], fused=True, st=sv_h)
                    free[0][slot].record(sv)
                    if peng.p2p is not None:
                        peng.dp_allreduce(c_pr, active, st=sp_h)
                    else:
                        with torch.cuda.stream(sp):
                            peng.dp_allreduce(c_pr, active)
                    peng.adam(B_glob, skel_nodes, active, 40.0, self.clip_first_step_only, world, plog[step], fused=True, st=sp_h)
                    free[1][slot].record(sp)
                elif world > 1 and self.overlap_nets:
                    # phases interleaved so both nets' collectives are issued back to back (NCCL runs them in
                    # issue order) while their compute overlaps on the two streams
                    with torch.cuda.stream(sv):
                        veng.dp_gather(arena, vrun, world)
                    with torch.cuda.stream(sp):
                        peng.dp_gather(arena, pruns, active, world)
                    with torch.cuda.stream(sv):
                        veng.dp_allgather(comm)
                    with torch.cuda.stream(sp):
                        peng.dp_allgather(active, comm)
                    with torch.cuda.stream(sv):
                        veng.dp_compute(arena, vrun, B_glob, world)
                    with torch.cuda.stream(sp):
                        peng.dp_compute(arena, pruns, active, B_glob, self.clip_epsilon, world)
                    with torch.cuda.stream(sv):
                        veng.dp_allreduce(comm)
                        veng.adam(B_glob, vlog[step])
                    with torch.cuda.stream(sp):
                        peng.dp_allreduce(comm, active)
                        peng.adam(B_glob, skel_nodes, active, 40.0, self.clip_first_step_only, world, plog[step])
                elif self.overlap_nets and self.prefetch_gather:
                    # gather + RunningNorm moments of minibatch k+1 run on the gather streams (second workspace) while
                    # minibatch k computes: they only read the arena, and they are 3 of the ~10 small kernels that
                    # otherwise sit between two minibatches on each net's critical path
                    slot = k & 1
                    if k == 0:
                        self._issue_prefetch(arena, plan, 0, 0, main)
                    if k + 1 < plan.n_mb:
                        self._issue_prefetch(arena, plan, k + 1, slot ^ 1, None)
                    gv, gp, ready, free = self._gather_ctx
                    # raw stream handles, no stream context managers: this loop issues ~16 000 launches per update and
                    # its Python glue was a third of the host's issue time
                    sv.wait_event(ready[0][slot])
                    veng.train_step_prefetched(arena, vrun, B_glob, vlog[step], slot, st=sv_h)
                    free[0][slot].record(sv)
                    sp.wait_event(ready[1][slot])
                    peng.train_step_prefetched(arena, pruns, B_glob, skel_nodes, active, self.clip_epsilon,
                                               40.0, self.clip_first_step_only, plog[step], slot, st=sp_h)
                    free[1][slot].record(sp)
                elif self.overlap_nets:
                    with torch.cuda.stream(sv):
                        veng.train_step(arena, vrun, B_glob, vlog[step], comm)
                    with torch.cuda.stream(sp):
                        peng.train_step(arena, pruns, B_glob, skel_nodes, active, self.clip_epsilon,
                                        40.0, self.clip_first_step_only, plog[step], comm)
                else:
                    veng.train_step(arena, vrun, B_glob, vlog[step], comm)
                    peng.train_step(arena, pruns, B_glob, skel_nodes, active, self.clip_epsilon,
                                    40.0, self.clip_first_step_only, plog[step], comm)
                step += 1
            if self.overlap_nets:
                # the next epoch's plan replaces the ordering tensors the in-flight kernels read
                plans_alive.append(plan)
        main.wait_stream(sv); main.wait_stream(sp)
        return self._finish_update(arena, plog, vlog, step, h2d, t_host0, cur_perm)

    def _finish_update(self, arena, plog, vlog, step, h2d, t_host0, cur_perm):
        host_enqueue_s = time.perf_counter() - t_host0        # host time to issue the whole update (no sync inside)
        logs = plog[:step].cpu().numpy()
        vlogs = vlog[:step].cpu().numpy()
        self.last_update_stats = {'h2d_bytes': arena.h2d_bytes + h2d, 'd2h_bytes': int(logs.nbytes + vlogs.nbytes),
                                  'host_enqueue_s': host_enqueue_s, 'steps': step, 'valid_steps': int(logs[:, 3].sum()) if step else 0,
                                  'value_loss': vlogs[:, 5].copy(), 'policy_log': logs}
        return self._summarise(arena, logs, cur_perm)

    def _update_two_threads(self, arena, perms, T, M, batch_design, n_mb, vlog, plog, main, sv, sp):
        """The minibatch loop of a single-GPU update with the critic issued from a second host thread.

        The two nets share nothing but the read-only arena and the epoch orderings, so every epoch's plan is built (and
        its ordering uploaded) first; then each net runs its own loop -- gather of minibatch k+1 on its gather stream,
        compute + Adam of minibatch k on its compute stream -- for all epochs.  Kernel order per stream, and therefore
        every result, is identical to the single-thread schedule."""
        import threading
        plans = []
        h2d = 0
        cur_perm = np.arange(T)
        for ep in range(self.opt_num_epochs):
            p = perms[ep] if perms is not None else self.draw_perm(T)          # agent :331-342
            cur_perm = cur_perm[p]
            plan = EpochPlan(arena, cur_perm, M, batch_design, n_mb=n_mb)
            cur_perm = plan.perm
            h2d += plan.h2d_bytes
            plans.append(plan)
        if self._gather_ctx is None or self._gather_ctx[0].device != self.device:
            ev = lambda: [[torch.cuda.Event() for _ in range(2)] for _ in range(2)]
            self._gather_ctx = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device), ev(), ev())
        gv, gp, ready, free = self._gather_ctx
        for st in (sv, sp, gv, gp):
            st.wait_stream(main)                     # the ordering uploads happened on the main stream
        veng, peng = self.value_net.engine, self.policy_net.engine
        clip_eps, first_only = self.clip_epsilon, self.clip_first_step_only
        errors = []

        def net_loop(net):
            try:
                with torch.cuda.device(self.device):
                    g, cs, eng = (gv, sv, veng) if net == 0 else (gp, sp, peng)
                    step = 0
                    for plan in plans:
                        runs = [plan.value_run(k) if net == 0 else plan.policy_runs(k) for k in range(plan.n_mb)]

                        def prefetch(k, slot):
                            g.wait_event(free[net][slot])          # the previous user of this workspace has finished
                            with torch.cuda.stream(g):
                                eng.prefetch(arena, runs[k], slot)
                                ready[net][slot].record(g)

                        for k in range(plan.n_mb):
                            slot = k & 1
                            if k == 0:
                                prefetch(0, 0)
                            if k + 1 < plan.n_mb:
                                prefetch(k + 1, slot ^ 1)
                            B_glob = int(plan.stage_graphs[k].sum())
                            with torch.cuda.stream(cs):
                                cs.wait_event(ready[net][slot])
                                if net == 0:
                                    eng.train_step_prefetched(arena, runs[k], B_glob, vlog[step], slot)
                                else:
                                    active = [bool(plan.stage_graphs[k, s] > 0) for s in range(3)]
                                    eng.train_step_prefetched(arena, runs[k], B_glob, int(plan.stage_nodes[k, 0]), active, clip_eps,
                                                              40.0, first_only, plog[step], slot)
                                free[net][slot].record(cs)
                            step += 1
            except BaseException as e:                  # noqa: BLE001  (re-raised on the calling thread)
                errors.append(e)

        th = threading.Thread(target=net_loop, args=(0,), name='sard-critic-issue')
        th.start()
        net_loop(1)
        th.join()
        if errors:
            raise errors[0]
        self._plans_alive = plans                        # the in-flight kernels read the ordering tensors
        return self.opt_num_epochs * n_mb, h2d, cur_perm

    def _issue_prefetch(self, arena: PackedRollout, plan, k: int, slot: int, after, comm=None, active=None):
        """GATHER phase of minibatch k of `plan` for both nets on their gather streams, into workspace `slot`."""
        if self._gather_ctx is None or self._gather_ctx[0].device != self.device:
            ev = lambda: [[torch.cuda.Event() for _ in range(2)] for _ in range(2)]
            self._gather_ctx = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device), ev(), ev())
        gv, gp, ready, free = self._gather_ctx
        veng, peng = self.value_net.engine, self.policy_net.engine
        for net, (g, eng, run) in enumerate(((gv, veng, plan.value_run(k)), (gp, peng, plan.policy_runs(k)))):
            if after is not None:
                g.wait_stream(after)                       # the epoch's ordering upload
            g.wait_event(free[net][slot])                  # the previous user of this workspace has finished
            if comm is None:                               # raw stream handle: no context switch on the host
                if net == 0:
                    eng.prefetch(arena, run, slot, st=g.cuda_stream)
                else:
                    eng.prefetch(arena, run, slot, st=g.cuda_stream)
                ready[net][slot].record(g)
                continue
            c = comm[net] if isinstance(comm, tuple) else comm
            if eng.board is not None:
                # moments all-gather over peer memory (one C call on the raw stream handle): no torch.distributed call and
                # no stream context on this path
                if net == 0:
                    eng.prefetch(arena, run, slot, c, st=g.cuda_stream)
                else:
                    eng.prefetch(arena, run, slot, c, active, st=g.cuda_stream)
                ready[net][slot].record(g)
                continue
            with torch.cuda.stream(g):
                if net == 0:
                    eng.prefetch(arena, run, slot, c)
                else:
                    eng.prefetch(arena, run, slot, c, active)
                ready[net][slot].record(g)

    def _summarise(self, arena: PackedRollout, logs: np.ndarray, last_perm: np.ndarray) -> dict:
        """Epoch log with the reference's keys (agent :371-387)."""
        log2 = dict()
        attr_ids = np.flatnonzero(arena.stage_np == 1)
        if len(attr_ids):
            num_agents = arena.num_nodes_np[attr_ids]
            log2['num_agent_mean'] = float(np.mean(num_agents))
            log2['num_agent_std'] = float(np.std(num_agents))
            # the packed host copies are still there (pinned staging of this arena): no device round trip for the log
            host = getattr(arena, '_host', None)
            edst = host['edst'].numpy() if host is not None else arena.edst.cpu().numpy()
            eptr = host['edge_ptr'].numpy() if host is not None else arena.edge_ptr.cpu().numpy()
            designs = Counter(tuple(edst[eptr[i]:eptr[i + 1]].tolist()) for i in attr_ids)
            log2['num_designs'] = len(designs)
            log2['num_designs_std'] = float(np.std(list(designs.values())))
            acts = host['actions'].numpy() if host is not None else arena.actions.cpu().numpy()
            a = [np.abs(acts[arena.node_ptr_np[i]:arena.node_ptr_np[i + 1], 1:3]).mean() for i in attr_ids]
            log2['attr_mean'] = float(np.mean(a))
            log2['attr_std'] = float(np.std(a))
        valid = logs[:, 3] > 0.5 if len(logs) else np.zeros(0, bool)
        for j, k in enumerate(('surr_loss', 'entropy', 'approx_kl')):
            v = logs[valid, j] if len(logs) else np.zeros(0)
            log2[k] = float(np.sum(v) / (np.count_nonzero(v) + 1e-6))
        return log2

    # ---- single-step API of the reference ------------------------------------------------------------------
    def update_value(self, states, returns):
        """agent_pg.py:18-25 for one list of states (one optimiser step)."""
        with torch.cuda.device(self.device):
            arena = PackedRollout.from_lists(states, device=self.device)
            self._prepare(arena)
            arena.returns.copy_(returns.reshape(-1).to(self.device, torch.float32))
            order = Ordering(arena, np.arange(arena.T))
            self.value_net.engine.train_step(arena, order.run(0, arena.T), arena.T, None, self.comm)

    def ppo_loss(self, states, actions, advantages, fixed_log_probs):
        """agent :390-409.  Forward + loss (+ backward into the engine's gradient bucket); returns
        ``(surr_loss tensor, log dict, valid_loss)``.  The returned tensor carries no autograd graph: where the
        reference calls ``surr_loss.backward(); clip_policy_grad(); optimizer_policy.step()`` (agent :354-357),
        call ``policy_step_from_loss()``."""
        with torch.cuda.device(self.device):
            arena = PackedRollout.from_lists(states, actions, device=self.device)
            self._prepare(arena)
            arena.advantages.copy_(advantages.reshape(-1).to(self.device, torch.float32))
            arena.fixed_log_probs.copy_(fixed_log_probs.reshape(-1).to(self.device, torch.float32))
            order = Ordering(arena, stage_partition(arena.stage_np))
            runs = order.stage_runs(0, arena.T)
            peng = self.policy_net.engine
            peng.GA.zero_()
            for s in (2, 1, 0):
                if runs[s] is not None:
                    peng.run_stage(s, arena, runs[s], 1, self.clip_epsilon, 1.0 / arena.T)
            stats = peng.GA[:8].cpu().numpy()
            n_skel = int(arena.num_nodes_np[arena.stage_np == 0].sum())
            surr = -float(stats[0]) / arena.T
            kl = float(stats[1]) / arena.T
            ent = float(stats[4]) + (float(stats[2]) / n_skel if n_skel else 0.0)
            log = {'surr_loss': surr, 'entropy': ent, 'approx_kl': kl}
            valid = not (kl > 0.1 or surr > 10)
            self._pending = (arena, runs)
            return torch.tensor(surr, device=self.device), log, valid

    def policy_step_from_loss(self):
        """Gradient clip + Adam (gated on device like agent :404-407) on the gradients the last ``ppo_loss()`` left in
        the policy engine's bucket: the ``backward / clip_policy_grad / optimizer_policy.step`` of agent :354-357."""
        if getattr(self, '_pending', None) is None:
            raise RuntimeError('policy_step_from_loss() needs a preceding ppo_loss()')
        arena, runs = self._pending
        self._pending = None
        with torch.cuda.device(self.device):
            n_skel = int(arena.num_nodes_np[arena.stage_np == 0].sum())
            self.policy_net.engine.adam(arena.T, n_skel, [r is not None for r in runs], 40.0, self.clip_first_step_only, 1)

    def clip_policy_grad(self):
        """Clipping is fused into the optimiser step (``sard_adam_prepare``); kept for API compatibility."""
        return None
