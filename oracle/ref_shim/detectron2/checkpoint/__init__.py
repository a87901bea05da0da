import os
import torch


class DetectionCheckpointer:
    """resume_or_load: loads `<save_dir>/model_final.pth` if present (state-dict under "model"),
    else leaves the random initialisation (detectron2 would load cfg.MODEL.WEIGHTS)."""

    def __init__(self, model, save_dir="", **kw):
        self.model = model
        self.save_dir = save_dir

    def resume_or_load(self, path, resume=True):
        f = os.path.join(self.save_dir, "model_final.pth")
        if os.path.isfile(f):
            sd = torch.load(f, map_location="cpu")
            self.model.load_state_dict(sd.get("model", sd), strict=False)
        return {}
