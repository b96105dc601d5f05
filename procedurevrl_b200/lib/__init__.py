"""Host-side mirror of the reference's `lib` package surface for the hot path:
lib.config.defaults.get_cfg, lib.models.build.{MODEL_REGISTRY, build_model},
lib.models.vit.{VisionTransformer, vit_base_patch16_224_develop}."""
