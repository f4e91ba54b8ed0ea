"""``BaseModel`` with the reference's checkpoint contract (look2hear/models/utils/base_model.py:34-94)."""
import torch
import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self, sample_rate, in_chan=1):
        super().__init__()
        self._sample_rate = sample_rate
        self._in_chan = in_chan

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def sample_rate(self):
        return self._sample_rate

    @staticmethod
    def load_state_dict_in_audio(model, pretrained_dict):
        """Load the ``audio_model.*`` entries of a Lightning checkpoint (base_model.py:48-57)."""
        model_dict = model.state_dict()
        update = {k[len("audio_model.") :]: v for k, v in pretrained_dict.items() if "audio_model" in k}
        model_dict.update(update)
        model.load_state_dict(model_dict)
        return model

    @staticmethod
    def from_pretrain(pretrained_model_conf_or_path, *args, **kwargs):
        """Instantiate from ``{"model_name", "state_dict", "model_args", "infos"}`` (base_model.py:60-69)."""
        from . import get

        # reference checkpoints hold plain python objects next to the tensors (torch 1.11 semantics)
        conf = torch.load(pretrained_model_conf_or_path, map_location="cpu", weights_only=False)
        model = get(conf["model_name"])(*args, **kwargs)
        model.load_state_dict(conf["state_dict"])
        return model

    def serialize(self):
        """Same dictionary as base_model.py:71-86; ``pytorch_lightning`` is optional here."""
        versions = dict(torch_version=str(torch.__version__))
        try:
            import pytorch_lightning as pl

            versions["pytorch_lightning_version"] = pl.__version__
        except ImportError:
            versions["pytorch_lightning_version"] = None
        return dict(
            model_name=self.__class__.__name__,
            state_dict=self.get_state_dict(),
            model_args=self.get_model_args(),
            infos=dict(software_versions=versions),
        )

    def get_state_dict(self):
        return self.state_dict()

    def get_model_args(self):
        raise NotImplementedError
