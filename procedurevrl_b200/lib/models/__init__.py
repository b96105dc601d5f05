from .build import MODEL_REGISTRY, build_model  # noqa: F401
from .vit import VisionTransformer, vit_base_patch16_224_develop  # noqa: F401
from .mvit import MViT, MViT_encoder  # noqa: F401,E402
