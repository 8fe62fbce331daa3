"""Batched ``select_action`` for rollout workers.

The reference moves the policy to the CPU and lets each of its sampler processes run batch-1 forwards
(``khrylib/rl/agents/agent.py:114-115``, ``design_opt/agents/transform2act_agent.py:64-100``).  Here the policy
stays on the GPU in the learner process; workers send their current state through a queue, the server gathers
whatever arrived within a short window (``max_batch`` states or ``max_wait_ms``), runs ONE batched
``Transform2ActPolicy.select_action`` (orbit tie and symmetrizer per graph) and hands each worker the rows of its
graph.  The call a worker makes keeps the reference's shape::

    action, action_env = client.select_action([state], mean_action)     # policy.py:339-381

Works with threads and with forked / spawned processes (``multiprocessing`` queues; states and actions are numpy).
Requests with different ``mean_action`` are never mixed in one batch.

A client may carry a ``key`` = the name of a design subgroup (``LearnedGroup.set_cur_subgroup_name``): its requests are
served with the policy's group switched to that subgroup (embedding, orbit tie and symmetrizer all follow it), so the
neighbour subgroups that ``optimize_policy`` evaluates (reference ``transform2act_agent.py:244-256``) can be sampled
concurrently through one server instead of one after the other.  Requests with different keys are never mixed.
"""
from __future__ import annotations

import multiprocessing as mp
import queue
import threading
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

_STOP = ('__stop__',)


def _to_numpy_state(state: Sequence) -> list:
    """Detach tensors so the state can cross a process boundary; the policy accepts numpy slots."""
    out = []
    for v in state:
        if hasattr(v, 'detach'):
            v = v.detach().cpu().numpy()
        out.append(np.asarray(v))
    return out


class ActionClient:
    """Worker-side handle.  One per worker; create it with :meth:`ActionServer.client` BEFORE forking."""

    def __init__(self, cid: int, request_q, response_q, key=None):
        self.cid, self._req, self._resp, self.key = cid, request_q, response_q, key

    def select_action(self, x: Sequence[Sequence], mean_action: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """``x``: list of states (the reference passes one) -> ``(action, action_env)`` float64 ``[sum N, 8]``."""
        states = [_to_numpy_state(s) for s in x]
        self._req.put((self.cid, (bool(mean_action), self.key), states))
        res = self._resp.get()
        if isinstance(res, BaseException):
            raise res
        return res


class ActionServer:
    """Owns the policy; batches the workers' requests.  ``policy`` needs ``select_action(states, mean_action)`` returning
    row-concatenated ``(action, action_env)`` for the list of states, like ``Transform2ActPolicy``."""

    def __init__(self, policy, max_batch: int = 64, max_wait_ms: float = 2.0, ctx: Optional[mp.context.BaseContext] = None):
        self.policy = policy
        self.max_batch, self.max_wait = int(max_batch), float(max_wait_ms) * 1e-3
        self._ctx = ctx if ctx is not None else mp.get_context()
        self._req = self._ctx.Queue()
        self._resp: List = []
        self._thread: Optional[threading.Thread] = None
        self.batches = 0            # statistics: batched forwards run / states served
        self.served = 0

    def client(self, key=None) -> ActionClient:
        q = self._ctx.Queue()
        self._resp.append(q)
        return ActionClient(len(self._resp) - 1, self._req, q, key)

    # -- serving ---------------------------------------------------------------------------------------------
    def start(self):
        self._thread = threading.Thread(target=self.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self._req.put(_STOP)
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def _collect(self):
        """One batch: block for the first request, then take what arrives within the window (same mean_action)."""
        first = self._req.get()
        if first == _STOP:
            return None, None
        batch, carry = [first], None
        n_states = len(first[2])
        deadline = time.monotonic() + self.max_wait
        while n_states < self.max_batch:
            left = deadline - time.monotonic()
            if left <= 0:
                break
            try:
                r = self._req.get(timeout=left)
            except queue.Empty:
                break
            if r == _STOP or r[1] != first[1]:
                carry = r                       # a stop, or a request of another sampling mode / subgroup: next round
                break
            batch.append(r)
            n_states += len(r[2])
        return batch, carry

    def serve_forever(self):
        carry = None
        while True:
            if carry is not None:
                if carry == _STOP:
                    return
                self._req.put(carry)             # back into the queue: order between different workers is free
                carry = None
            batch, carry = self._collect()
            if batch is None:
                return
            self._serve(batch)

    def _serve(self, batch):
        states, spans, n = [], [], 0
        for cid, _, sts in batch:
            rows = sum(int(np.asarray(s[3]).reshape(-1)[0]) for s in sts)       # slot 3: num_nodes
            states.extend(sts)
            spans.append((cid, n, n + rows))
            n += rows
        try:
            mean_action, key = batch[0][1]
            group = getattr(self.policy, 'group', None) if key is not None else None
            prev = group.cur_subgroup_name if group is not None else None      # (not store/restore: that single
            if group is not None:                                               #  backup slot belongs to optimize_policy)
                group.set_cur_subgroup_name(key)
            try:
                action, action_env = self.policy.select_action(states, mean_action)
            finally:
                if group is not None:
                    group.set_cur_subgroup_name(prev)
            for cid, a, b in spans:
                self._resp[cid].put((np.array(action[a:b]), np.array(action_env[a:b])))
        except BaseException as e:               # noqa: BLE001  -- the worker re-raises it
            for cid, _, _ in spans:
                self._resp[cid].put(RuntimeError(f'ActionServer: {type(e).__name__}: {e}'))
        self.batches += 1
        self.served += len(states)
