"""TEST INFRASTRUCTURE -- plain PyTorch fp32 restatement of the DS-RNN policy
forward (a floating-point path, so the oracle is a torch fp32 reference).

Follows, for one rollout step (infer=True, seq_len 1):
  SRNN.forward                pytorchBaselines/a2c_ppo_acktr/srnn_model.py:409-504
  HumanHumanEdgeRNN.forward   srnn_model.py:201-215  (+ RNNBase._forward_gru :37-50)
  EdgeAttention.forward       srnn_model.py:256-339
  HumanNodeRNN.forward        srnn_model.py:149-173
  DiagGaussian.forward        pytorchBaselines/a2c_ppo_acktr/distributions.py:85-94

Pinned against the reference's own `Policy.act(deterministic=True)` run in the
build container on both shipped checkpoints (oracle/gen_golden_dsrnn.py ->
tests/golden/dsrnn_*.npz).  Takes a reference-layout state_dict (name ->
tensor); never imports the product.
"""
import math

import torch


def _gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def forward(sd, robot_node, temporal_edges, spatial_edges, h_node, h_edge, masks):
    """Returns dict(value[N,1], action_mean[N,2], actor_features[N,256], h_node[N,1,128], h_edge[N,H+1,256])."""
    p = lambda k: sd[k].to(torch.float32)
    N, H = spatial_edges.shape[0], spatial_edges.shape[1]
    m = masks.reshape(N, 1, 1).to(torch.float32)
    robot_node = robot_node.reshape(N, 7).to(torch.float32)
    t_in = temporal_edges.reshape(N, 2).to(torch.float32)
    s_in = spatial_edges.reshape(N, H, 2).to(torch.float32)
    h_edge = h_edge.reshape(N, H + 1, 256).to(torch.float32) * m
    h_node = h_node.reshape(N, 1, 128).to(torch.float32) * m

    def edge(prefix, x, h):
        e = torch.relu(x @ p(prefix + ".encoder_linear.weight").t() + p(prefix + ".encoder_linear.bias"))
        return _gru_cell(e, h, p(prefix + ".gru.weight_ih_l0"), p(prefix + ".gru.weight_hh_l0"),
                         p(prefix + ".gru.bias_ih_l0"), p(prefix + ".gru.bias_hh_l0"))

    o_t = edge("base.humanhumanEdgeRNN_temporal", t_in, h_edge[:, 0])                       # [N,256]
    o_s = edge("base.humanhumanEdgeRNN_spatial", s_in.reshape(N * H, 2),
               h_edge[:, 1:].reshape(N * H, 256)).reshape(N, H, 256)                        # [N,H,256]

    q = o_t @ p("base.attn.temporal_edge_layer.0.weight").t() + p("base.attn.temporal_edge_layer.0.bias")
    k = o_s @ p("base.attn.spatial_edge_layer.0.weight").t() + p("base.attn.spatial_edge_layer.0.bias")
    score = (q.unsqueeze(1) * k).sum(-1) * (H / math.sqrt(64.0))                            # [N,H]
    alpha = torch.softmax(score, dim=-1)
    c = (alpha.unsqueeze(-1) * o_s).sum(1)                                                  # [N,256]

    rob = robot_node @ p("base.robot_linear.weight").t() + p("base.robot_linear.bias")      # [N,3]
    enc = torch.relu(rob @ p("base.humanNodeRNN.encoder_linear.weight").t() + p("base.humanNodeRNN.encoder_linear.bias"))
    emb = torch.relu(torch.cat([o_t, c], -1) @ p("base.humanNodeRNN.edge_attention_embed.weight").t()
                     + p("base.humanNodeRNN.edge_attention_embed.bias"))
    x = torch.cat([enc, emb], -1)                                                           # [N,128]
    h_n = _gru_cell(x, h_node[:, 0], p("base.humanNodeRNN.gru.weight_ih_l0"), p("base.humanNodeRNN.gru.weight_hh_l0"),
                    p("base.humanNodeRNN.gru.bias_ih_l0"), p("base.humanNodeRNN.gru.bias_hh_l0"))
    y = h_n @ p("base.humanNodeRNN.output_linear.weight").t() + p("base.humanNodeRNN.output_linear.bias")

    def mlp(prefix, v):
        v = torch.tanh(v @ p(prefix + ".0.weight").t() + p(prefix + ".0.bias"))
        return torch.tanh(v @ p(prefix + ".2.weight").t() + p(prefix + ".2.bias"))

    actor = mlp("base.actor", y)
    critic = mlp("base.critic", y)
    value = critic @ p("base.critic_linear.weight").t() + p("base.critic_linear.bias")
    mean = actor @ p("dist.fc_mean.weight").t() + p("dist.fc_mean.bias")
    return dict(value=value, action_mean=mean, actor_features=actor, h_node=h_n.unsqueeze(1),
                h_edge=torch.cat([o_t.unsqueeze(1), o_s], 1), attention=alpha)


def random_state_dict(seed=0, scale=1.0):
    """Random-init weights of the DS-RNN architecture in the reference's state_dict layout
    (orthogonal-ish scale; srnn_model.py:28-32, 378-402).  For synthetic benchmarks."""
    g = torch.Generator().manual_seed(seed)
    shapes = {
        "encoder_linear.weight": (64, 2), "encoder_linear.bias": (64,),
        "gru.weight_ih_l0": (768, 64), "gru.weight_hh_l0": (768, 256), "gru.bias_ih_l0": (768,), "gru.bias_hh_l0": (768,),
    }
    sd = {}
    for pre in ("base.humanhumanEdgeRNN_temporal", "base.humanhumanEdgeRNN_spatial"):
        for k, s in shapes.items():
            sd[pre + "." + k] = s
    sd.update({
        "base.attn.temporal_edge_layer.0.weight": (64, 256), "base.attn.temporal_edge_layer.0.bias": (64,),
        "base.attn.spatial_edge_layer.0.weight": (64, 256), "base.attn.spatial_edge_layer.0.bias": (64,),
        "base.robot_linear.weight": (3, 7), "base.robot_linear.bias": (3,),
        "base.humanNodeRNN.encoder_linear.weight": (64, 3), "base.humanNodeRNN.encoder_linear.bias": (64,),
        "base.humanNodeRNN.edge_embed.weight": (64, 256), "base.humanNodeRNN.edge_embed.bias": (64,),
        "base.humanNodeRNN.edge_attention_embed.weight": (64, 512), "base.humanNodeRNN.edge_attention_embed.bias": (64,),
        "base.humanNodeRNN.gru.weight_ih_l0": (384, 128), "base.humanNodeRNN.gru.weight_hh_l0": (384, 128),
        "base.humanNodeRNN.gru.bias_ih_l0": (384,), "base.humanNodeRNN.gru.bias_hh_l0": (384,),
        "base.humanNodeRNN.output_linear.weight": (256, 128), "base.humanNodeRNN.output_linear.bias": (256,),
        "base.actor.0.weight": (256, 256), "base.actor.0.bias": (256,),
        "base.actor.2.weight": (256, 256), "base.actor.2.bias": (256,),
        "base.critic.0.weight": (256, 256), "base.critic.0.bias": (256,),
        "base.critic.2.weight": (256, 256), "base.critic.2.bias": (256,),
        "base.critic_linear.weight": (1, 256), "base.critic_linear.bias": (1,),
        "base.human_node_final_linear.weight": (2, 256), "base.human_node_final_linear.bias": (2,),
        "dist.fc_mean.weight": (2, 256), "dist.fc_mean.bias": (2,), "dist.logstd._bias": (2, 1),
    })
    out = {}
    for k, s in sd.items():
        if len(s) == 2 and not k.endswith("_bias"):
            out[k] = torch.randn(s, generator=g) * (scale / math.sqrt(s[1]))
        else:
            out[k] = torch.randn(s, generator=g) * 0.05
    return out
