"""gnn_fluid_dynamics_b200 - B200-native (sm_100a) message-passing processor for the FVGN / MGN /
Flux / Conservative / VertPot models of aj-dray/gnn-fluid-dynamics.

Importing the models requires the in-tree CUDA library (``lib/libgnnfd_b200.so``); there is no CPU
or PyTorch fallback for the hot path.  ``graph`` / ``mesh`` are importable without it.
"""
__version__ = "0.1.0"
