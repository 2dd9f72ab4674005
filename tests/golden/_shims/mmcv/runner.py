import os

import torch


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    if not os.path.exists(filename):      # all checkpoints are stripped from the reference tree
        return None
    sd = torch.load(filename, map_location=map_location)
    sd = sd.get("state_dict", sd)
    model.load_state_dict(sd, strict=strict)
    return sd
