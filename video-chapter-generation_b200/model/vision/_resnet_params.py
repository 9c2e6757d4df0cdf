"""Parameter containers with the state-dict layout of torchvision's ResNet-50 (v1.5 bottlenecks).

They only HOLD parameters (same names, shapes and default initialisation as torchvision 0.26 ``resnet50``:
kaiming-normal convs, unit BatchNorm); all arithmetic happens in libvcg_b200.so.  BatchNorm is deliberately NOT an
``nn.BatchNorm2d`` subclass: reference caller #1 nulls the running statistics of every ``nn.BatchNorm2d`` it finds
(test_video_segment_point.py:116-122); the engine folds eval-mode statistics into the convolutions, so the running
statistics must survive that loop (DESIGN.md "BatchNorm mode").  What that loop does to the reference —
batch statistics at inference — is available as the opt-in mode ``TwoStream.bn_batch_stats`` (vcg_b200/bn_batch.py).
"""
import torch
import torch.nn as nn


class ConvParams(nn.Module):
    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.stride = cin, cout, (k, k), (stride, stride)
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        raise RuntimeError("this module only holds parameters; run the model through TwoStream.forward")


class BatchNormStats(nn.Module):
    """Eval-mode BatchNorm2d parameters + running statistics (folded into the preceding conv by the engine)."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = c, eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        raise RuntimeError("this module only holds parameters; run the model through TwoStream.forward")


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        self.conv1 = ConvParams(inplanes, planes, 1)
        self.bn1 = BatchNormStats(planes)
        self.conv2 = ConvParams(planes, planes, 3, stride)   # v1.5: the stride sits on the 3x3
        self.bn2 = BatchNormStats(planes)
        self.conv3 = ConvParams(planes, planes * 4, 1)
        self.bn3 = BatchNormStats(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(ConvParams(inplanes, planes * 4, 1, stride), BatchNormStats(planes * 4))
        self.stride = stride


class FcParams(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)


class ResNet50Params(nn.Module):
    """conv1 / bn1 / layer1..4 / fc, keyed exactly like torchvision.models.resnet50().state_dict()."""

    def __init__(self):
        super().__init__()
        self.conv1 = ConvParams(3, 64, 7, 2)
        self.bn1 = BatchNormStats(64)
        inplanes = 64
        for idx, (blocks, planes) in enumerate(zip((3, 4, 6, 3), (64, 128, 256, 512)), start=1):
            layer = []
            for i in range(blocks):
                stride = 2 if (i == 0 and idx > 1) else 1
                layer.append(Bottleneck(inplanes, planes, stride, downsample=(i == 0)))
                inplanes = planes * 4
            setattr(self, f"layer{idx}", nn.Sequential(*layer))
        self.fc = FcParams(2048, 1000)

    def forward(self, x):
        raise RuntimeError("the ResNet-50 backbone runs inside libvcg_b200.so; call TwoStream.forward")


def _weights_version(module):
    v = 0
    for t in list(module.parameters()) + list(module.buffers()):
        v += t._version + (t.data_ptr() % 1000003)
    return v


def _unimodal_vision_forward(module, x, shift_div):
    """Shared body of Resnet50TSM.forward / Resnet50.forward (--data_mode image): the backbone and the
    nn.Linear(T*2048, 2) head run inside libvcg_b200.so (VCG_MODALITY_VISION engine).  No CPU / eager fallback."""
    import os
    import torch
    if module.head is None:
        raise RuntimeError("call build_chapter_head() first")
    if not x.is_cuda:
        raise RuntimeError("image-only scoring needs CUDA inputs: the B200 implementation has no CPU fallback")
    if module.training:
        raise RuntimeError("inference-only: call .eval() first")
    from vcg_b200.engine import Engine
    precision = getattr(module, "precision", os.environ.get("VCG_PRECISION", "bf16"))
    chunk = getattr(module, "vision_chunk", int(os.environ.get("VCG_VISION_CHUNK", "32")))
    key = (str(x.device), precision, chunk, _weights_version(module))
    if getattr(module, "_engine", None) is None or module._engine_key != key:
        if getattr(module, "_engine", None) is not None:
            module._engine.close()
        eng = Engine(module.segments_size, "mlp", precision, True, 128, chunk, 128, shift_div, device=x.device,
                     modality="vision")
        eng.load_state_dict(module.state_dict())
        object.__setattr__(module, "_engine", eng)
        object.__setattr__(module, "_engine_key", key)
    with torch.no_grad():
        return module._engine.forward_vision(x)
