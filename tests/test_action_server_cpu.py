"""CPU: the batching / routing logic of the rollout-side action server with a stand-in policy (the GPU test runs it
with the real Transform2ActPolicy)."""
import multiprocessing as mp
import threading

import numpy as np

from sard_b200.action_server import ActionServer


class FakePolicy:
    """action row i of a state = [state id, node index, mean flag, ...]; records the batch sizes it saw."""

    def __init__(self):
        self.batch_sizes = []

    def select_action(self, states, mean_action):
        self.batch_sizes.append(len(states))
        rows = []
        for s in states:
            n = int(np.asarray(s[3]).reshape(-1)[0])
            sid = float(np.asarray(s[0])[0, 0])
            for j in range(n):
                rows.append([sid, j, float(mean_action)] + [0.0] * 5)
        a = np.asarray(rows, dtype=np.float64)
        return a, a + 100.0


def make_state(sid, n):
    obs = np.full((n, 4), float(sid))
    return [obs, np.zeros((2, 2 * (n - 1)), dtype=np.int64), np.array([2]), np.array([n]), np.arange(n), np.zeros(1), np.zeros(1), np.zeros(1)]


def _worker(client, wid, n_steps, out_q):
    ok = True
    for t in range(n_steps):
        n = 3 + (wid + t) % 5
        sid = wid * 1000 + t
        a, ae = client.select_action([make_state(sid, n)], mean_action=bool(wid % 2))
        ok &= a.shape == (n, 8) and np.all(a[:, 0] == sid) and np.all(a[:, 1] == np.arange(n)) and np.all(a[:, 2] == wid % 2)
        ok &= bool(np.all(ae == a + 100.0))
    out_q.put((wid, bool(ok)))


def test_threads_get_their_own_rows_and_requests_are_batched():
    pol = FakePolicy()
    srv = ActionServer(pol, max_batch=8, max_wait_ms=20.0)
    clients = [srv.client() for _ in range(6)]
    srv.start()
    out = mp.get_context().Queue()
    th = [threading.Thread(target=_worker, args=(c, i, 25, out)) for i, c in enumerate(clients)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    srv.stop()
    res = dict(out.get() for _ in th)
    assert all(res.values()) and len(res) == 6
    assert srv.served == 6 * 25
    assert max(pol.batch_sizes) > 1, 'concurrent requests must be merged into batches'
    assert max(pol.batch_sizes) <= 8


def test_forked_processes():
    ctx = mp.get_context('fork')
    pol = FakePolicy()
    srv = ActionServer(pol, max_batch=16, max_wait_ms=10.0, ctx=ctx)
    clients = [srv.client() for _ in range(3)]
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(c, i, 10, out)) for i, c in enumerate(clients)]
    srv.start()
    for p in procs:
        p.start()
    res = dict(out.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    srv.stop()
    assert all(res.values()) and srv.served == 30


def test_policy_errors_reach_the_worker():
    class Broken:
        def select_action(self, states, mean_action):
            raise ValueError('boom')
    srv = ActionServer(Broken(), max_batch=4, max_wait_ms=1.0)
    c = srv.client()
    srv.start()
    try:
        c.select_action([make_state(1, 3)])
        raised = False
    except RuntimeError as e:
        raised = 'boom' in str(e)
    srv.stop()
    assert raised
