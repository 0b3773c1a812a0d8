"""Shared by the golden generator (reference Python) and the test (mirror): model, data, evaluation callback and cases of the
AutoQuant check."""
import torch


class Net(torch.nn.Module):
    """fx-traceable, with a foldable Conv -> BN pair, a BN -> Conv pair that cannot fold forward (padding) and a Linear."""

    def __init__(self):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.bn1 = torch.nn.BatchNorm2d(8)
        self.relu1 = torch.nn.ReLU()
        self.conv2 = torch.nn.Conv2d(8, 8, 3, padding=1, bias=False)
        self.bn2 = torch.nn.BatchNorm2d(8)
        self.relu2 = torch.nn.ReLU()
        self.pool = torch.nn.AdaptiveAvgPool2d(2)
        self.flatten = torch.nn.Flatten()
        self.fc = torch.nn.Linear(32, 5)

    def forward(self, x):
        x = self.relu1(self.bn1(self.conv1(x)))
        x = self.relu2(self.bn2(self.conv2(x)))
        return self.fc(self.flatten(self.pool(x)))


def make_model():
    torch.manual_seed(0)
    m = Net()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for bn in (m.bn1, m.bn2):     # batch norms with real statistics and a wide spread of scales
            bn.weight.copy_(torch.rand(8, generator=g) * 3 + 0.2)
            bn.bias.copy_(torch.randn(8, generator=g) * 0.3)
            bn.running_mean.copy_(torch.randn(8, generator=g) * 0.2)
            bn.running_var.copy_(torch.rand(8, generator=g) * 2 + 0.1)
    return m.eval()


def make_loader():
    """A real DataLoader (the reference insists on the type) over 64 seeded images, batches of 16, no shuffling."""
    g = torch.Generator().manual_seed(7)
    return torch.utils.data.DataLoader(torch.randn(64, 3, 12, 12, generator=g), batch_size=16)


def make_eval_callback(fp32_model, loader):
    """Score = -mean squared difference to the fp32 model's outputs over the loader (0 for the fp32 model itself)."""
    with torch.no_grad():
        targets = [fp32_model(b) for b in loader]

    def eval_callback(model, _num_samples=None):
        device = next(model.parameters()).device
        with torch.no_grad():
            err = sum(float(((model(b.to(device)) - t.to(device)) ** 2).mean()) for b, t in zip(loader, targets))
        return -err / len(loader)
    return eval_callback


ADAROUND_ITERATIONS = 32
# name -> (param_bw, output_bw, allowed_accuracy_drop); the score is -MSE against the fp32 outputs, so fp32 scores 0
CASES = {"w8a8_loose": (8, 8, 1.0),          # batch-norm folding already meets the target
         "w4a8_tight": (4, 8, 1e-4),         # W32 passes, folding does not: every stage runs, the best result is returned
         "w8a16_tight": (8, 16, 1e-7),       # output_bw >= 16 narrows the quant-scheme candidates to activation tf
         "w8a8_impossible": (8, 8, 0.0)}     # even W32 misses the target: early exit, everything None
