"""Duck-typed `Boxes` / `Instances` exposing exactly the members the reference's consumers touch
(`instances_to_json`, reference src/probabilistic_inference/inference_utils.py:454-502, and
`src/apply_net.py:91-96`): len(), .pred_boxes.tensor, .scores, .pred_classes, .pred_cls_probs,
.pred_boxes_covariance, .has(name), .image_size.  If detectron2 is importable the real classes are
used instead, so the objects are interchangeable inside the original harness."""
import torch

try:  # pragma: no cover - detectron2 is not installed in the build image
    from detectron2.structures import Boxes, Instances  # noqa: F401
    HAVE_DETECTRON2 = True
except Exception:  # noqa: BLE001
    HAVE_DETECTRON2 = False

    class Boxes:
        def __init__(self, tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape((-1, 4))
            assert tensor.dim() == 2 and tensor.size(-1) == 4
            self.tensor = tensor

        def __len__(self):
            return self.tensor.shape[0]

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            return Boxes(self.tensor[item])

        def to(self, device):
            return Boxes(self.tensor.to(device))

        def area(self):
            b = self.tensor
            return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

        @property
        def device(self):
            return self.tensor.device

    class Instances:
        def __init__(self, image_size, **kwargs):
            self._image_size = image_size
            self._fields = {}
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith("_"):
                super().__setattr__(name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            if name == "_fields" or name not in self._fields:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
            return self._fields[name]

        def set(self, name, value):
            if len(self._fields):
                assert len(self) == len(value)
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def get(self, name):
            return self._fields[name]

        def get_fields(self):
            return self._fields

        def to(self, *args, **kwargs):
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v.to(*args, **kwargs) if hasattr(v, "to") else v)
            return ret

        def __getitem__(self, item):
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v[item])
            return ret

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            raise NotImplementedError("Empty Instances does not support __len__!")
