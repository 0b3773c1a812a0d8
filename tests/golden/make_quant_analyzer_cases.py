"""Shared by make_quant_analyzer_golden.py (runs the REFERENCE's QuantAnalyzer) and the tests (run aimet_b200's): the
model, the data, the callbacks and the cases. The model has no batch norm: the reference's analyzer folds batch norms
first, which is outside the hot path (and needs its full native library)."""
import torch

# name: (quantsim config, quant scheme, ignore a module?)
CASES = {
    "net_default_tfe": (None, "tf_enhanced", False),
    "net_perchannel_tfe": ("default_config_per_channel.json", "tf_enhanced", False),
    "net_default_tf": (None, "tf", False),
    "net_ignore_tfe": (None, "tf_enhanced", True),
}


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.relu1 = torch.nn.ReLU()
        self.conv2 = torch.nn.Conv2d(8, 16, 3, stride=2, padding=1)
        self.relu2 = torch.nn.ReLU()
        self.conv3 = torch.nn.Conv2d(16, 16, 3, padding=1)
        self.relu3 = torch.nn.ReLU()
        self.pool = torch.nn.AdaptiveAvgPool2d(1)
        self.flat = torch.nn.Flatten()
        self.fc = torch.nn.Linear(16, 10)

    def forward(self, x):
        x = self.relu1(self.conv1(x))
        x = self.relu2(self.conv2(x))
        x = self.relu3(self.conv3(x))
        return self.fc(self.flat(self.pool(x)))


def make_model():
    torch.manual_seed(0)
    return Net().eval()


def make_data():
    torch.manual_seed(1)
    batches = [torch.randn(4, 3, 16, 16) * (1.0 + 0.25 * i) for i in range(3)]
    target = torch.randn(4, 10)
    return batches, target


def callbacks(batches, target):
    def forward_pass(model, _):
        with torch.no_grad():
            for x in batches:
                model(x.to(next(model.parameters()).device))

    def evaluate(model, _):
        with torch.no_grad():
            dev = next(model.parameters()).device
            return float(-torch.nn.functional.mse_loss(model(batches[0].to(dev)), target.to(dev)))

    return forward_pass, evaluate
