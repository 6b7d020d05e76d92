"""rl_b200.objectives -- only the value-estimator corner of ``torchrl.objectives`` is on the hot path."""
