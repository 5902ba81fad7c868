"""Synthetic local-BA problems: the generator lives in orb_slam3_rgbl_b200.synthetic (bench.py uses it too)."""
from orb_slam3_rgbl_b200.synthetic import _quat_yaw, _rot, make_ba_problem as make_problem, ba_args as args  # noqa: F401
