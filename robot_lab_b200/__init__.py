"""robot_lab_b200 - B200-native per-step MDP pipeline behind robot_lab's term / manager / env API."""

__version__ = "0.1.0"
