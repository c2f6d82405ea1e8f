"""The backbone checker lives with the other oracles (oracle/torch_backbone.py); kept importable under its old name."""
from oracle.torch_backbone import run_graph  # noqa: F401
