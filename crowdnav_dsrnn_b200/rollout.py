"""CUDA-graph rollout: one deterministic rollout step (`Policy.act` -> `CrowdVecEnv.step`, i.e. train.py:243-261 /
evaluation.py:119-134 without the host round trips) captured once and replayed.

At small batches (BASELINE.json configs[0]/[1]: 16 / 1024 envs) a step is ~15 kernel launches of a few microseconds
each, so the eager loop is bound by launch latency; a graph replay submits the whole step with one call.  Two graphs are
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
        if hx is not None:
            self.sets[cur]["h_node"].copy_(hx["human_node_rnn"].reshape(n, 1, 128))
            self.sets[cur]["h_edge"].copy_(hx["human_human_edge_rnn"].reshape(n, H + 1, 256))
        if masks is not None:
            self.sets[cur]["masks"].copy_(masks.reshape(n, 1))
        assert obs["robot_node"].data_ptr() == eng.bufs[cur].robot_node.data_ptr(), "obs must be the engine's current buffer"
        self.parity = cur                 # which set holds the inputs of the next step
        self.graphs = [None, None]
        self.steps = 0
        self.launches_per_step = 0
        # eager warm-up of both parities (lazy handle / workspace creation must not happen inside a capture)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
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
        src = eng.bufs[p]
        with torch.no_grad():
            self.policy.cuda_forward(src.obs(), {"human_node_rnn": a["h_node"], "human_human_edge_rnn": a["h_edge"]},
                                     a["masks"], need_features=False,
                                     out=dict(h_node=b["h_node"], h_edge=b["h_edge"], value=a["value"], mean=a["mean"]))
            dst = eng.step(a["mean"], auto_reset=True)            # flips eng.cur to p ^ 1 and writes bufs[p ^ 1]
            torch.sub(1.0, dst.done.to(torch.float32).unsqueeze(1), out=b["masks"])
        return dst

    def step(self):
        """Replay one rollout step; returns the StepBuffers holding its outputs (valid until the step after next)."""
        p = self.parity
        self.graphs[p].replay()
        self.parity = p ^ 1
        self.eng.cur = self.parity
        self.steps += 1
        return self.eng.bufs[self.parity]

    def hidden(self):
        s = self.sets[self.parity]
        return {"human_node_rnn": s["h_node"], "human_human_edge_rnn": s["h_edge"]}, s["masks"]
