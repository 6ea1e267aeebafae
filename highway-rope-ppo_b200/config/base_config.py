"""highway-v0 configuration used by every experiment.

Same keys and values as the reference's ``config/base_config.py:5-39`` (callers pass this dict
to ``make_env`` unchanged); assembled from parts so that the observation block can be reused
by the vectorised benchmarks.
"""


def _symmetric(limit):
    return [-limit, limit]


OBSERVATION_CONFIG = dict(
    type="Kinematics",
    vehicles_count=15,                      # rows of the observation, ego included
    features=["x", "y", "vx", "vy"],
    normalize=True,
    features_range=dict(
        x=_symmetric(100), y=_symmetric(100),
        vx=_symmetric(30), vy=_symmetric(30),
        presence=[0, 1], cos_h=_symmetric(1), sin_h=_symmetric(1),
    ),
    absolute=False,
    order="sorted",
)

ACTION_CONFIG = dict(type="ContinuousAction", longitudinal=True, lateral=True)

HIGHWAY_CONFIG = dict(
    observation=OBSERVATION_CONFIG,
    action=ACTION_CONFIG,
    simulation_frequency=15,
    policy_frequency=1,
    duration=40,
    lanes_count=4,
    vehicles_count=50,
    vehicles_density=2,
    collision_reward=-1,
    right_lane_reward=0.1,
    high_speed_reward=0.4,
    lane_change_reward=-0.05,               # dead key in highway-v0 1.10.1, kept for parity
    reward_speed_range=[20, 30],
)
