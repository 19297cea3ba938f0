"""CUDA-graph rollout: one deterministic rollout step (`CrowdVecEnv.step` -> `Policy.act` on the new observation, i.e. the
loop of train.py:243-261 / evaluation.py:119-134 without the host round trips) captured once and replayed.

At small batches (BASELINE.json configs[0]/[1]: 16 / 1024 envs) a step is 7 kernel launches of a few microseconds
each (plus the host-side bookkeeping of `act`), so the eager loop is bound by launch latency; a graph replay submits the whole step with one call.  Two graphs are
captured because every buffer is ping-ponged (the engine double-buffers its outputs, the hidden state and masks
alternate between two static sets): graph 0 reads set 0 and writes set 1, graph 1 the reverse.
"""
import torch


class GraphedRollout(object):
    def __init__(self, policy, venv, obs, hx=None, masks=None):
        """`obs` must be the observation returned by the engine's LAST reset()/step() (it lives in its current buffer)."""
        self.policy, self.venv = policy, venv
        eng = venv.engine
        dev, n, H = eng.device, eng.n, eng.h
        self.eng = eng
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.sets = [dict(h_node=z(n, 1, 128), h_edge=z(n, H + 1, 256), masks=z(n, 1), value=z(n, 1), mean=z(n, 2)) for _ in range(2)]
        cur = eng.cur
        assert obs["robot_node"].data_ptr() == eng.bufs[cur].robot_node.data_ptr(), "obs must be the engine's current buffer"
        hx0 = {"human_node_rnn": z(n, 1, 128) if hx is None else hx["human_node_rnn"].reshape(n, 1, 128),
               "human_human_edge_rnn": z(n, H + 1, 256) if hx is None else hx["human_human_edge_rnn"].reshape(n, H + 1, 256)}
        m0 = z(n, 1) if masks is None else masks.reshape(n, 1)
        self.parity = cur                 # graph p steps the env out of buffer p (action = sets[p]["mean"]) into buffer p ^ 1
        self.graphs = [None, None]
        self.steps = 0
        self.launches_per_step = 0
        # A captured step is: crowd step (+ swap-in of spare episodes, fork of their refill) -> masks -> forward on the new
        # observation -> join of the refill, so the spawn of the next episodes runs beside the forward.  The forward of the
        # FIRST observation therefore happens here, eagerly; it leaves the action in sets[cur] and the hidden state in
        # sets[cur ^ 1], where graph `cur` expects them.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            a, b = self.sets[cur], self.sets[cur ^ 1]
            eng.join()
            b["masks"].copy_(m0)
            policy.cuda_forward(obs, hx0, m0, need_features=False,
                                out=dict(h_node=b["h_node"], h_edge=b["h_edge"], value=a["value"], mean=a["mean"]))
            eng.join()
            policy.start_refill_of(eng)        # from here on the forward starts the refill of the step before it
            # eager warm-up of both parities (lazy handle / workspace creation must not happen inside a capture)
            l0 = eng.launches + policy.gpu_launches
            self._one_step(self.parity)
            self.launches_per_step = eng.launches + policy.gpu_launches - l0
            self._one_step(self.parity ^ 1)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in (self.parity, self.parity ^ 1):        # engine.cur == p when graph p is captured
            assert eng.cur == p
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(p)
            self.graphs[p] = g
        # the two captures executed nothing; the state is where the warm-up left it (two steps in)
        self.steps = 2

    def _one_step(self, p):
        eng, a, b = self.eng, self.sets[p], self.sets[p ^ 1]
        with torch.no_grad():
            dst = eng.step(a["mean"], auto_reset=True, defer_refill=True)   # flips eng.cur to p ^ 1, writes bufs[p ^ 1]
            b["masks"] = dst.not_done                             # 1 - done, written by the step kernel itself
            self.policy.cuda_forward(dst.obs(), {"human_node_rnn": b["h_node"], "human_human_edge_rnn": b["h_edge"]},
                                     b["masks"], need_features=False,
                                     out=dict(h_node=a["h_node"], h_edge=a["h_edge"], value=b["value"], mean=b["mean"]))
            eng.join()
        return dst

    def step(self):
        """Replay one rollout step; returns the StepBuffers holding its outputs (valid until the step after next)."""
        p = self.parity
        self.graphs[p].replay()
        self.parity = p ^ 1
        self.eng.cur = self.parity
        self.steps += 1
        return self.eng.bufs[self.parity]

    def close(self):
        """Detach the forward from the env: eager `act` / `step` calls behave as usual afterwards."""
        self.policy.start_refill_of(None)
        self.eng.join()

    def hidden(self):
        """(hidden state, masks) that belong to the observation of the last step() -- what the next `act` would be given."""
        s = self.sets[self.parity]
        return {"human_node_rnn": s["h_node"], "human_human_edge_rnn": s["h_edge"]}, s["masks"]
