"""Rollout-side policy throughput: per-state select_action calls (what each reference sampler does, here on the GPU)
vs the batching ActionServer with W worker threads.  Diagnostic; full derl.yml nets, point_nav shape."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sard_b200 import synthetic  # noqa: E402
from sard_b200.action_server import ActionServer  # noqa: E402
from sard_b200.config import derl_config  # noqa: E402
from sard_b200.transform2act_agent import Transform2ActAgent  # noqa: E402


def main():
    W = int(os.environ.get('WORKERS', 20))          # derl.yml num_threads
    S = int(os.environ.get('STATES', 2000))
    shape = synthetic.ENV_SHAPES['point_nav']
    torch.manual_seed(0)
    ag = Transform2ActAgent(derl_config(), torch.float32, 'cuda:0', 0, 1, env=synthetic.SyntheticEnv(shape))
    roll = synthetic.generate_rollout(shape, S, seed=2, min_exec=20, max_exec=60)
    states = roll.to_traj_batch().states
    ag.policy_net.eval()
    for s in states[:20]:
        ag.policy_net.select_action([s], False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in states[:400]:
        ag.policy_net.select_action([s], False)
    torch.cuda.synchronize()
    per_state = 400 / (time.perf_counter() - t0)
    srv = ActionServer(ag.policy_net, max_batch=W, max_wait_ms=1.0)
    clients = [srv.client() for _ in range(W)]

    def work(w):
        for i in range(w, len(states), W):
            clients[w].select_action([states[i]], False)

    srv.start()
    th = [threading.Thread(target=work, args=(w,)) for w in range(W)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    srv.stop()
    print(f'per-state calls: {per_state:.0f} actions/s; action server with {W} workers: {len(states) / dt:.0f} actions/s '
          f'({srv.batches} batched forwards for {srv.served} states, mean batch {srv.served / max(srv.batches, 1):.1f})')


if __name__ == '__main__':
    main()
