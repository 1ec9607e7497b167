"""B200-native rollout + NAF hot path with the robotic_manipulator_rloa API."""
from .rl_framework import ManipulatorFramework  # noqa: F401
