"""Shared by the golden generator (reference Python) and the test (mirror): model, data and cases of the AdaRound check."""
import torch

# name -> (quantsim config, weight bitwidth, iterations per layer)
CASES = {"default_bw4": ("default", 4, 48), "per_channel_bw8": ("per_channel", 8, 32)}


def make_model():
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 8, 3, padding=1),
                               torch.nn.BatchNorm2d(8), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten(),
                               torch.nn.Linear(32, 5))


def make_batches():
    g = torch.Generator().manual_seed(7)
    return [torch.randn(16, 3, 12, 12, generator=g) for _ in range(4)]
