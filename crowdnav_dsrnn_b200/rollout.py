"""CUDA-graph rollout: one deterministic rollout step (`CrowdVecEnv.step` -> `Policy.act` on the new observation, i.e. the
loop of train.py:243-261 / evaluation.py:119-134 without the host round trips) captured once and replayed.

At small batches (BASELINE.json configs[0]/[1]: 16 / 1024 envs) a step is 7 kernel launches of a few microseconds
each (plus the host-side bookkeeping of `act`), so the eager loop is bound by launch latency; a graph replay submits the whole step with one call.  Two graphs are
captured because every buffer is ping-ponged (the engine double-buffers its outputs, the hidden state and masks
alternate between two static sets): graph 0 reads set 0 and writes set 1, graph 1 the reverse.
"""
import torch


class GraphedRollout(object):
    def __init__(self, policy, venv, obs, hx=None, masks=None, edge_image=True):
        """`obs` must be the observation returned by the engine's LAST reset()/step() (it lives in its current buffer)."""
        self.policy, self.venv = policy, venv
        eng = venv.engine
        dev, n, H = eng.device, eng.n, eng.h
        self.eng = eng
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.sets = [dict(h_node=z(n, 1, 128), h_edge=z(n, H + 1, 256), masks=z(n, 1), value=z(n, 1), mean=z(n, 2)) for _ in range(2)]
        # resident split-bf16 image of each edge-state buffer (bf16x3 forward only): the forward that writes h_edge of a set also
        # writes its image, the next forward reads its A operand from the image by TMA instead of converting the fp32 state
        self.use_image = bool(edge_image) and policy.precision == "bf16x3"
        self.image_mode = edge_image if edge_image in ("in", "out") else "both"      # "in" / "out": timing experiments only (stale images)
        if self.use_image:
            for s_ in self.sets:
                s_["image"] = tuple(torch.zeros(n * (H + 1), 256, dtype=torch.bfloat16, device=dev) for _ in range(2))
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
            policy.cuda_forward(obs, hx0, m0, need_features=False,       # (creates the library handle)
                                out=dict(h_node=b["h_node"], h_edge=b["h_edge"], value=a["value"], mean=a["mean"]))
            if self.use_image:       # once more, now also writing the image of b's edge state (same inputs, same outputs)
                policy.set_edge_image(None, b["image"])
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
            if self.use_image:       # this forward reads the image the previous one wrote for b's state and writes a's
                self.policy.set_edge_image(None if self.image_mode == "out" else b["image"],
                                           None if self.image_mode == "in" else a["image"])
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

    def step_eager(self):
        """The same step launched eagerly (same kernels, same schedule incl. the refill forked beside the attention kernel):
        what bench.py uses to time individual kernels with CUDA events, which cannot be read back from a graph replay."""
        p = self.parity
        dst = self._one_step(p)
        self.parity = p ^ 1
        self.steps += 1
        return dst

    def close(self):
        """Detach the forward from the env: eager `act` / `step` calls behave as usual afterwards."""
        self.policy.start_refill_of(None)
        if self.use_image:
            self.policy.set_edge_image(None, None)
        self.eng.join()

    def hidden(self):
        """(hidden state, masks) that belong to the observation of the last step() -- what the next `act` would be given."""
        s = self.sets[self.parity]
        return {"human_node_rnn": s["h_node"], "human_human_edge_rnn": s["h_edge"]}, s["masks"]


class PipelinedRollout(object):
    """The same rollout loop for ONE batch of envs cut into two independent halves that are software-pipelined on two
    streams inside one CUDA graph.

    A rollout step is two machine-filling kernels (crowd step: issue-bound, edge GRUs: tensor-bound) followed by a tail of
    small dependent kernels (folded projection, attention, node RNN + heads: ~0.17 ms at 16384 envs, most SMs idle).  Envs
    are independent, so half B's crowd step does not have to wait for half A's tail.  One replay runs

        stream 0:  edge A -> [projection, attention, node/heads A] ------------------> crowd step A
        stream 1:            (edge A done) crowd step B -> edge B -> [projection, attention, node/heads B]
                                                                      (edge B done) ----^

    i.e. the big kernels form one chain edge A -> step B -> edge B -> step A and each tail runs beside the other half's
    crowd step.  Every env still does exactly one `act` and one `step` per replay; the halves are skewed by half a
    period: after a replay half A holds a fresh observation (its forward is the first thing of the next replay) and half
    B holds observation AND action.  Results are those of the plain loop bit for bit: the env RNG is keyed by the global
    env id (`env_id_offset`), and no kernel reads across envs (tests/test_gpu_rollout_graph.py).
    """

    def __init__(self, policy, config, n_envs, device, seed=0, phase="train", env_id_offset=0, nenv=None, split=None):
        from .engine import CrowdEngine

        dev = torch.device(device)
        self.policy, self.device = policy, dev
        n, H = int(n_envs), int(config.sim.human_num)
        n_a = int(split) if split else self.default_split(n, H, dev)
        if not 0 < n_a < n:
            raise ValueError("split must leave both halves non-empty")
        self.sizes = (n_a, n - n_a)
        total = n if nenv is None else nenv
        self.engines = [CrowdEngine(config, self.sizes[0], dev, phase=phase, seed=seed, env_id_offset=env_id_offset, nenv=total),
                        CrowdEngine(config, self.sizes[1], dev, phase=phase, seed=seed, env_id_offset=env_id_offset + n_a, nenv=total)]
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.sets = [[dict(h_node=z(m, 1, 128), h_edge=z(m, H + 1, 256), value=z(m, 1), mean=z(m, 2)) for _ in range(2)]
                     for m in self.sizes]
        self.side = torch.cuda.Stream(device=dev, priority=-1)
        self.edge_done = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in self.edge_done:                  # torch creates the cudaEvent_t lazily; the library needs the handle
            ev.record()
            assert ev.cuda_event, "torch.cuda.Event has no handle after record()"
        self.graphs = [None, None]
        self.parity = 0
        self.steps = 0
        lib_bytes = None
        with torch.no_grad():
            for eng in self.engines:
                buf = eng.reset()
                buf.not_done.zero_()               # masks of the very first forward (train.py:196-204: zeros)
            # lazily created things (library handle, weight images) must exist before a capture
            a = self.engines[0]
            policy.cuda_forward(a.bufs[a.cur].obs(), {"human_node_rnn": z(self.sizes[0], 1, 128), "human_human_edge_rnn": z(self.sizes[0], H + 1, 256)},
                                z(self.sizes[0], 1), need_features=False)
            from . import _lib
            lib_bytes = [_lib.load().cn_dsrnn_workspace_bytes(m, H) for m in self.sizes]
            self.workspaces = [torch.empty(b, dtype=torch.uint8, device=dev) for b in lib_bytes]
            # half B enters the loop with its first action computed (its crowd step is the first thing it does in a replay)
            self._forward(1, 0)
            for eng in self.engines:
                eng.join()
            l0 = self._launch_count()
            self._cycle(0)                         # eager warm-up of both parities
            self.launches_per_step = self._launch_count() - l0
            self._cycle(1)
        torch.cuda.synchronize(dev)
        for p in (0, 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._cycle(p)
            self.graphs[p] = g
        policy.start_refill_of(None)
        policy.set_edge_event(None)
        self.steps = 2

    @staticmethod
    def default_split(n, H, dev):
        """First half = a whole number of waves of the edge kernel (one 256-row tile per CTA pair and wave)."""
        pairs = max(1, torch.cuda.get_device_properties(dev).multi_processor_count // 2)
        rows_per_wave = 256 * pairs
        waves = max(1, round(n * (H + 1) / rows_per_wave / 2))
        n_a = (waves * rows_per_wave) // (H + 1)
        return n_a if 0 < n_a < n else n // 2

    def _launch_count(self):
        return self.policy.gpu_launches + sum(e.launches for e in self.engines)

    def _forward(self, i, q):
        """Forward of half i on its current observation: hidden sets[i][q] -> sets[i][q ^ 1], value / action into sets[i][q]."""
        eng, src, dst = self.engines[i], self.sets[i][q], self.sets[i][q ^ 1]
        buf = eng.bufs[eng.cur]
        self.policy.start_refill_of(eng)           # host-side hooks, read by the forward call below
        self.policy.set_edge_event(self.edge_done[i])
        self.policy.cuda_forward(buf.obs(), {"human_node_rnn": src["h_node"], "human_human_edge_rnn": src["h_edge"]},
                                 buf.not_done, need_features=False, workspace=self.workspaces[i],
                                 out=dict(h_node=dst["h_node"], h_edge=dst["h_edge"], value=src["value"], mean=src["mean"]))

    def _cycle(self, p):
        """forward A, step B, forward B, step A (see the class docstring); runs eagerly or under capture.  The forwards are
        launched on a HIGH-priority stream: the CTAs of a tail kernel are then dispatched ahead of the queued CTAs of the
        other half's crowd step whenever an SM frees a slot (a graph kernel node keeps its stream's priority)."""
        main = torch.cuda.current_stream(self.device)
        hi = self.side
        a, b = self.engines
        with torch.no_grad():
            hi.wait_stream(main)
            with torch.cuda.stream(hi):
                self._forward(0, p)                                    # records edge_done[0] behind edge A
            main.wait_event(self.edge_done[0])
            b.step(self.sets[1][p]["mean"], auto_reset=True, defer_refill=True)
            hi.wait_stream(main)
            with torch.cuda.stream(hi):
                self._forward(1, p ^ 1)                                # records edge_done[1] behind edge B
                b.join()
            main.wait_event(self.edge_done[1])                         # tail A precedes edge B on `hi`: done as well
            a.step(self.sets[0][p]["mean"], auto_reset=True, defer_refill=True)
            main.wait_stream(hi)

    def step(self):
        """Replay one period: one `act` + one `step` for every env."""
        p = self.parity
        self.graphs[p].replay()
        self.parity = p ^ 1
        for e in self.engines:
            e.cur ^= 1
        self.steps += 1

    def state(self):
        """Per half: its engine, the StepBuffers of its last crowd step and the hidden state its NEXT forward reads.  Half 1
        is half a period ahead (class docstring): its hidden state already includes the forward of that observation, whose
        `action` / `value` are returned with it."""
        p = self.parity                 # parity of the next replay
        a, b = self.engines
        out = [dict(engine=a, buffers=a.bufs[a.cur], hidden={"human_node_rnn": self.sets[0][p]["h_node"],
                                                              "human_human_edge_rnn": self.sets[0][p]["h_edge"]})]
        # the last replay (parity p ^ 1) ran forward(1, p): hidden sets[1][p] -> sets[1][p ^ 1], action / value into sets[1][p]
        out.append(dict(engine=b, buffers=b.bufs[b.cur], action=self.sets[1][p]["mean"], value=self.sets[1][p]["value"],
                        hidden={"human_node_rnn": self.sets[1][p ^ 1]["h_node"], "human_human_edge_rnn": self.sets[1][p ^ 1]["h_edge"]}))
        return out

    def close(self):
        for e in self.engines:
            e.join()
            e.close()
