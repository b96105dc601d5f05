from .defaults import CfgNode, get_cfg  # noqa: F401
