"""Model registry / DDP boundary (reference lib/models/build.py:8-54).

`MODEL_REGISTRY.get(cfg.MODEL.MODEL_NAME)(cfg)` builds the module, `build_model` moves it to the current CUDA
device and, for NUM_GPUS > 1, wraps it in DistributedDataParallel: one process per GPU, a single bucketed NCCL
gradient all-reduce over NVLink/NVSwitch overlapped with backward -- the only collective on
the path.  `find_unused_parameters=True` as in the reference (build.py:49-53): its fine-tuning flows depend on it --
`construct_optimizer` freezes the encoder for TRAIN.LINEAR only AFTER `build_model` has wrapped the module, and the
NUM_SEG > 0 forecasting path never touches `order_tfm.pad_embedding` -- so parameters registered as trainable may see
no gradient in a step."""
import torch


class Registry:
    """The two calls of fvcore.common.registry.Registry the reference uses: register() and get()."""

    def __init__(self, name):
        self._name, self._map = name, {}

    def register(self, obj=None):
        def deco(o):
            self._map[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self._map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._map[name]

    def __contains__(self, name):
        return name in self._map


MODEL_REGISTRY = Registry("MODEL")


def build_model(cfg, gpu_id=None):
    """build.py:17-54: construct cfg.MODEL.MODEL_NAME, move to GPU, wrap in DDP when NUM_GPUS > 1."""
    if not torch.cuda.is_available():
        raise RuntimeError("build_model: the sm_100a implementation needs a CUDA device (there is no CPU fallback)")
    assert cfg.NUM_GPUS <= torch.cuda.device_count(), "Cannot use more GPU devices than available"
    name = cfg.MODEL.MODEL_NAME
    model = MODEL_REGISTRY.get(name)(cfg)
    cur_device = torch.cuda.current_device() if gpu_id is None else gpu_id
    model = model.cuda(device=cur_device)
    if cfg.NUM_GPUS > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
        model = wrap_data_parallel(model, cfg, cur_device)
    return model


def wrap_data_parallel(model, cfg, device=None):
    """The data-parallel boundary of build.py:49-53: replicas + one bucketed gradient all-reduce (NCCL on GPUs;
    the CPU tests drive the same wrapper over gloo)."""
    bucket_mb = cfg.B200.GRAD_BUCKET_MB if "B200" in cfg else 64
    ids = None if device is None else [device]
    return torch.nn.parallel.DistributedDataParallel(
        module=model, device_ids=ids, output_device=device, find_unused_parameters=True,
        gradient_as_bucket_view=True, bucket_cap_mb=bucket_mb, broadcast_buffers=False)
