"""ResNet student wrapper: LT/_models/torchvision/resnet.py:21-47 (forward_features -> layer4 map, forward_pool -> avgpool).
The convolutional student runs on cuDNN through torchvision / torch.autograd (library code, stated as such in DESIGN.md);
this package's kernels cover the teacher and the loss of the distillation step."""
from __future__ import annotations

from typing import Dict

from torch import Tensor, nn


class ResNetModelWrapper(nn.Module):
    def __init__(self, model: nn.Module) -> None:
        super().__init__()
        self._model = model
        self._feature_dim: int = model.fc.in_features

    def feature_dim(self) -> int:
        return self._feature_dim

    def forward_features(self, x: Tensor) -> Dict[str, Tensor]:
        m = self._model
        x = m.maxpool(m.relu(m.bn1(m.conv1(x))))
        x = m.layer4(m.layer3(m.layer2(m.layer1(x))))
        return {"features": x}

    def forward_pool(self, x: Dict[str, Tensor]) -> Dict[str, Tensor]:
        return {"pooled_features": self._model.avgpool(x["features"])}

    def get_model(self) -> nn.Module:
        return self._model


class EmbeddingModel(nn.Module):
    """LT/_models/embedding_model.py surface the method reads: `.wrapped_model`, `.embed_dim`."""

    def __init__(self, wrapped_model: nn.Module) -> None:
        super().__init__()
        self.wrapped_model = wrapped_model
        self.embed_dim = wrapped_model.feature_dim()
