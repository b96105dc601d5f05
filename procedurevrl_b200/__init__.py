"""procedurevrl_b200: Blackwell-native (sm_100a) implementation of the ProcedureVRL TimeSformer hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/pvrl.h), ops.py (ctypes binding), engine.py (the explicit
forward/backward schedule of the divided space-time encoder), lib/ (host-side mirror of the reference's
lib.models / lib.config surface: MODEL_REGISTRY, build_model, vit_base_patch16_224_develop, get_cfg)."""
__version__ = "0.1.0"
