import torch


def _get_device(mod) -> torch.device:
    if isinstance(mod, torch.nn.Module):
        for t in list(mod.parameters()) + list(mod.buffers()):
            return t.device
    return torch.device("cpu")
