"""Rollout storage for the DS-RNN PPO update.

Same interface and buffer semantics as the reference's `SRNNRolloutStorage`
(pytorchBaselines/a2c_ppo_acktr/storage.py:14-292): `obs[key][T+1, N, ...]`, `recurrent_hidden_states[key][T+1, N, ...]`,
`rewards / value_preds / returns / action_log_probs / actions / masks / bad_masks`, `insert`, `after_update`,
`compute_returns`, `recurrent_generator`.  What differs is how it is laid out for a GPU rollout of 10^4..10^5 envs:

* everything is allocated once on the rollout device (the reference allocates on the CPU and `.to(device)`s),
  `insert` is a handful of `copy_`s on the current stream and never synchronises;
* the recurrent generator only ever reads hidden-state slot 0 of a chunk (storage.py:251-253), so
  `keep_hidden_history=False` (default for large N) keeps two slots -- [0] the chunk's initial state, [-1] the running
  state -- instead of T+1 (T=30, N=16384, H=20: 0.35 GB instead of 10.6 GB for the edge states alone);
* minibatches are gathered with one `index_select` per buffer instead of a Python loop over envs
  (storage.py:240-262), and the returns recursion is one fused reverse loop over [N] vectors.
"""
import torch


class SRNNRolloutStorage:
    def __init__(self, num_steps, num_processes, obs_shape, action_space, human_node_rnn_size, human_human_edge_rnn_size,
                 recurrent_cell_type="GRU", device="cpu", keep_hidden_history=True):
        if recurrent_cell_type != "GRU":
            raise NotImplementedError("only GRU cells are on the DS-RNN path (srnn_model.py:27-33)")
        T, N = int(num_steps), int(num_processes)
        opts = dict(dtype=torch.float32, device=torch.device(device))
        self.obs = {k: torch.zeros(T + 1, N, *obs_shape[k].shape, **opts) for k in obs_shape}
        self.human_num = obs_shape["spatial_edges"].shape[0]
        self.keep_hidden_history = bool(keep_hidden_history)
        slots = T + 1 if self.keep_hidden_history else 2
        self.recurrent_hidden_states = {
            "human_node_rnn": torch.zeros(slots, N, 1, human_node_rnn_size, **opts),
            "human_human_edge_rnn": torch.zeros(slots, N, self.human_num + 1, human_human_edge_rnn_size, **opts),
        }
        self.rewards = torch.zeros(T, N, 1, **opts)
        self.value_preds = torch.zeros(T + 1, N, 1, **opts)
        self.returns = torch.zeros(T + 1, N, 1, **opts)
        self.action_log_probs = torch.zeros(T, N, 1, **opts)
        if action_space.__class__.__name__ != "Box":
            raise NotImplementedError("only Box action spaces are supported")
        self.actions = torch.zeros(T, N, action_space.shape[0], **opts)
        self.masks = torch.ones(T + 1, N, 1, **opts)
        self.bad_masks = torch.ones(T + 1, N, 1, **opts)     # 0 where an episode ended on a time limit (storage.py:69-71)
        self.num_steps = T
        self.num_processes = N
        self.step = 0

    # ------------------------------------------------------------------ reference interface
    def to(self, device):
        for d in (self.obs, self.recurrent_hidden_states):
            for k in d:
                d[k] = d[k].to(device)
        for name in ("rewards", "value_preds", "returns", "action_log_probs", "actions", "masks", "bad_masks"):
            setattr(self, name, getattr(self, name).to(device))
        return self

    @property
    def device(self):
        return self.rewards.device

    def hidden_at(self, step):
        """Hidden states the policy starts step `step` from (train.py:233-237 reads `recurrent_hidden_states[key][step]`)."""
        if self.keep_hidden_history:
            return {k: v[step] for k, v in self.recurrent_hidden_states.items()}
        slot = 0 if step == 0 else 1          # [0] = state the chunk started from, [1] = state after the latest insert
        return {k: v[slot] for k, v in self.recurrent_hidden_states.items()}

    def obs_at(self, step):
        return {k: v[step] for k, v in self.obs.items()}

    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks, bad_masks=None):
        s = self.step
        for k in self.obs:
            self.obs[k][s + 1].copy_(obs[k].view_as(self.obs[k][s + 1]))
        slot = s + 1 if self.keep_hidden_history else 1
        for k, v in recurrent_hidden_states.items():
            self.recurrent_hidden_states[k][slot].copy_(v.view_as(self.recurrent_hidden_states[k][slot]))
        self.actions[s].copy_(actions)
        self.action_log_probs[s].copy_(action_log_probs.view_as(self.action_log_probs[s]))
        self.value_preds[s].copy_(value_preds.view_as(self.value_preds[s]))
        self.rewards[s].copy_(rewards.view_as(self.rewards[s]))
        self.masks[s + 1].copy_(masks.view_as(self.masks[s + 1]))
        if bad_masks is None:
            self.bad_masks[s + 1].fill_(1.0)
        else:
            self.bad_masks[s + 1].copy_(bad_masks.view_as(self.bad_masks[s + 1]))
        self.step = (s + 1) % self.num_steps

    def after_update(self):
        for k in self.obs:
            self.obs[k][0].copy_(self.obs[k][-1])
        for k in self.recurrent_hidden_states:
            self.recurrent_hidden_states[k][0].copy_(self.recurrent_hidden_states[k][-1])
        self.masks[0].copy_(self.masks[-1])
        self.bad_masks[0].copy_(self.bad_masks[-1])

    def compute_returns(self, next_value, use_gae, gamma, gae_lambda, use_proper_time_limits=True):
        """storage.py:132-176, all four branches, as one reverse loop."""
        T = self.num_steps
        rew, val, msk, bad, ret = self.rewards, self.value_preds, self.masks, self.bad_masks, self.returns
        if use_gae:
            val[-1] = next_value.view_as(val[-1])
            gae = torch.zeros_like(val[-1])
            for t in reversed(range(T)):
                delta = rew[t] + gamma * val[t + 1] * msk[t + 1] - val[t]
                gae = delta + gamma * gae_lambda * msk[t + 1] * gae
                if use_proper_time_limits:
                    gae = gae * bad[t + 1]
                ret[t] = gae + val[t]
        else:
            ret[-1] = next_value.view_as(ret[-1])
            for t in reversed(range(T)):
                r = ret[t + 1] * gamma * msk[t + 1] + rew[t]
                if use_proper_time_limits:
                    r = r * bad[t + 1] + (1 - bad[t + 1]) * val[t]
                ret[t] = r

    def recurrent_generator(self, advantages, num_mini_batch, perm=None, env_slice=None):
        """Yields the reference's 8-tuple per minibatch (storage.py:222-292): [T*n, ...] time-major chunks of n =
        N // num_mini_batch whole env trajectories and their slot-0 hidden states.

        `perm` overrides the env permutation (default `torch.randperm(N)` on the CPU generator, exactly what the reference
        draws, so the same `torch.manual_seed` reproduces its minibatches).  `env_slice=(a, b)` restricts every minibatch
        to positions [a, b) of its n envs -- gradient accumulation over sub-chunks for rollouts whose activations do not
        fit at once (ppo.py `max_envs_per_pass`)."""
        T, N = self.num_steps, self.num_processes
        if N < num_mini_batch:
            raise AssertionError("PPO requires the number of processes (%d) to be greater than or equal to the number of "
                                 "PPO mini batches (%d)." % (N, num_mini_batch))
        n = N // num_mini_batch
        if perm is None:
            perm = torch.randperm(N)
        perm = torch.as_tensor(perm, dtype=torch.int64).to(self.device)
        for start in range(0, n * num_mini_batch, n):      # N % num_mini_batch envs are left out, as in PPO.update
            ind = perm[start:start + n]
            if env_slice is not None:
                ind = ind[env_slice[0]:env_slice[1]]
            yield self.gather(ind, advantages)

    def gather(self, ind, advantages):
        T, k = self.num_steps, ind.numel()
        take = lambda buf: buf[:T].index_select(1, ind).reshape(T * k, *buf.shape[2:])
        obs_batch = {key: take(v) for key, v in self.obs.items()}
        hx_batch = {key: v[0].index_select(0, ind) for key, v in self.recurrent_hidden_states.items()}
        return (obs_batch, hx_batch, take(self.actions), take(self.value_preds), take(self.returns), take(self.masks),
                take(self.action_log_probs), None if advantages is None else take(advantages))
