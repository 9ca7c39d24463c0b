"""Minimal runtime shared by the trainer-side engines and the agent-side savers
(the subset of dlrover/python/common the Flash Checkpoint path needs)."""
