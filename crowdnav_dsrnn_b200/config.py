"""`Config`: the flag tree of crowd_nav/configs/config.py:9-214, same attribute
names, defaults and derived values, so code written against the reference's
`config.<section>.<name>` keeps working.  Any object with this attribute tree
(including the reference's own `Config`) is accepted by this package.

Unlike the reference (class attributes shared by every instance) each
`Config()` owns its sections, and the values the reference derives at class
creation time (config.py:37-40, 51-54, 66-74, 87-92) are re-derived by
`derive()` from the primary switches passed to the constructor.
"""
from types import SimpleNamespace as _NS

TRAFFIC_SCENARIOS = ("circle_crossing", "square_crossing", "parallel_traffic", "perpendicular_traffic")
SIDE_PREF_SCENARIOS = ("side_pref_passing", "side_pref_overtaking", "side_pref_crossing")


class BaseConfig(_NS):
    """Attribute bag (same name as the reference's section type, config.py:4-6)."""


class Config(object):
    def __init__(self, social_metrics=False, train_val_sim=None, test_sim=None, time_step=0.25,
                 normalize_reward=False, kinematics="holonomic", human_num=None):
        self.test = BaseConfig(social_metrics=bool(social_metrics))
        self.sim = BaseConfig(
            render=False,
            train_val_sim=list(TRAFFIC_SCENARIOS if train_val_sim is None else train_val_sim),
            test_sim=list(TRAFFIC_SCENARIOS if test_sim is None else test_sim),
            square_width=20,
            group_human=False,
        )
        self.env = BaseConfig(env_name="CrowdSimDict-v0", time_limit=50, time_step=time_step, val_size=100,
                              test_size=500, randomize_attributes=True, seed=0)
        self.reward = BaseConfig(time_factor=False, normalize=bool(normalize_reward), potential_based=True,
                                 exponential=False, norm_zones=False, discomfort_dist_front=0.25,
                                 discomfort_dist_back=0.25, exp_denom=6, gamma=0.99, norm_zone_side="lhs",
                                 norm_zone_penalty=-0.5)
        self.humans = BaseConfig(visible=True, policy="orca", radius=0.3, v_pref=1, sensor="coordinates", FOV=2.0,
                                 goal_change_chance=0.25, end_goal_change_chance=1.0, random_radii=False,
                                 random_v_pref=False, random_unobservability=False, unobservable_chance=0.3,
                                 random_policy_changing=False)
        self.robot = BaseConfig(visible=False, policy="srnn", radius=0.3, v_pref=1, sensor="coordinates", FOV=2.0)
        self.noise = BaseConfig(add_noise=False, type="uniform", magnitude=0.1)
        self.lidar = BaseConfig(enable=False, viz=False,
                                cfg={"max_range": 5, "num_beams": 180, "robot_radius": self.robot.radius})
        self.action_space = BaseConfig(kinematics=kinematics)
        self.orca = BaseConfig(neighbor_dist=10, safety_space=0.15, time_horizon=5, time_horizon_obst=5)
        self.sf = BaseConfig(A=2.0, B=1, KI=1)
        self.social = BaseConfig(min_personal_space=0.2, max_walking_speed=1.5)
        self.ppo = BaseConfig(num_mini_batch=2, num_steps=30, recurrent_policy=True, epoch=5, clip_param=0.2,
                              value_loss_coef=0.5, entropy_coef=0.0, use_gae=True, gae_lambda=0.95)
        self.ConvGRU = BaseConfig(input_size=256, hidden_size=256)
        self.SRNN = BaseConfig(human_node_rnn_size=128, human_human_edge_rnn_size=256, human_node_input_size=3,
                               human_human_edge_input_size=2, human_node_output_size=256,
                               human_node_embedding_size=64, human_human_edge_embedding_size=64, attention_size=64)
        self.training = BaseConfig(lr=4e-5, eps=1e-5, alpha=0.99, max_grad_norm=0.5, num_env_steps=10e6,
                                   use_linear_lr_decay=False, save_interval=200, log_interval=20,
                                   use_proper_time_limits=False, cuda_deterministic=False, cuda=True,
                                   num_processes=12, output_dir="data/dummy", resume=False, load_path=None,
                                   overwrite=True, num_threads=1)
        self.derive()
        if human_num is not None:
            self.sim.human_num = int(human_num)

    def derive(self):
        """Recompute every value the reference derives from another flag."""
        test, sim, env, reward, humans = self.test, self.sim, self.env, self.reward, self.humans
        test.side_preference = any("side_pref" in s for s in sim.test_sim)
        special = test.social_metrics or test.side_preference
        sim.circle_radius = 4 if special else 6
        sim.human_num = 1 if test.side_preference else 5
        env.test_size = 2000 if test.social_metrics else (200 if test.side_preference else 500)
        assert reward.potential_based != reward.exponential
        norm = reward.normalize
        reward.success_reward = 1 if norm else 10
        reward.collision_penalty = -1 if norm else -20
        reward.timeout_penalty = -1 if norm else -20
        reward.discomfort_penalty_factor = (0.5 if norm else 10) * env.time_step
        reward.potential_factor = 0.1 if norm else 2
        reward.exp_factor = 0.025 if norm else 0.5
        humans.random_goal_changing = not test.side_preference
        humans.end_goal_changing = not test.side_preference
        return self
