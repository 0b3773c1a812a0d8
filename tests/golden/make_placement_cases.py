"""Shared by the golden generator (reference Python, ConnectedGraph) and the test (mirror, torch.fx): architectures whose
quantizer PLACEMENT -- which input / output / parameter quantizers exist and are enabled after QuantizationSimModel(...) --
is compared. No calibration: placement is decided by the graph and the config file alone."""
import torch
import torchvision


class TransformerBlock(torch.nn.Module):
    """LayerNorm / Linear / GELU / softmax / matmul / residual adds, written with modules and functionals."""

    def __init__(self, d=32, heads=4):
        super().__init__()
        self.heads = heads
        self.ln1 = torch.nn.LayerNorm(d)
        self.q = torch.nn.Linear(d, d)
        self.k = torch.nn.Linear(d, d)
        self.v = torch.nn.Linear(d, d)
        self.softmax = torch.nn.Softmax(dim=-1)
        self.proj = torch.nn.Linear(d, d)
        self.ln2 = torch.nn.LayerNorm(d)
        self.fc1 = torch.nn.Linear(d, 4 * d)
        self.act = torch.nn.GELU()
        self.fc2 = torch.nn.Linear(4 * d, d)

    def forward(self, x):
        h = self.ln1(x)
        q, k, v = self.q(h), self.k(h), self.v(h)
        att = self.softmax(torch.matmul(q, k.transpose(1, 2)) * 0.25)
        x = x + self.proj(torch.matmul(att, v))
        return x + self.fc2(self.act(self.fc1(self.ln2(x))))


class ConcatNet(torch.nn.Module):
    """Two branches joined by torch.cat, a depthwise convolution, PReLU, a max pool feeding two consumers, a Linear head."""

    def __init__(self):
        super().__init__()
        self.stem = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.bn = torch.nn.BatchNorm2d(8)
        self.act = torch.nn.PReLU()
        self.pool = torch.nn.MaxPool2d(2)
        self.left = torch.nn.Conv2d(8, 8, 1)
        self.right_dw = torch.nn.Conv2d(8, 8, 3, padding=1, groups=8)
        self.right_relu = torch.nn.ReLU6()
        self.merge = torch.nn.Conv2d(16, 8, 1)
        self.sig = torch.nn.Sigmoid()
        self.gap = torch.nn.AdaptiveAvgPool2d(1)
        self.flat = torch.nn.Flatten()
        self.fc = torch.nn.Linear(8, 4)

    def forward(self, x):
        x = self.pool(self.act(self.bn(self.stem(x))))
        y = torch.cat([self.left(x), self.right_relu(self.right_dw(x))], dim=1)
        y = self.merge(y)
        y = y * self.sig(y)
        return self.fc(self.flat(self.gap(y)))


def _tv(name, **kw):
    return lambda: getattr(torchvision.models, name)(**kw)


# name -> (constructor, config: "default" | "per_channel", input shape)
CASES = {
    "vgg11_bn": (_tv("vgg11_bn"), "default", (1, 3, 64, 64)),
    "alexnet": (_tv("alexnet"), "default", (1, 3, 96, 96)),
    "squeezenet1_1": (_tv("squeezenet1_1"), "default", (1, 3, 64, 64)),
    "densenet121": (_tv("densenet121"), "default", (1, 3, 64, 64)),
    "mnasnet0_5": (_tv("mnasnet0_5"), "default", (1, 3, 64, 64)),
    "googlenet": (_tv("googlenet", aux_logits=False, init_weights=False), "default", (1, 3, 64, 64)),
    "resnext50_32x4d": (_tv("resnext50_32x4d"), "per_channel", (1, 3, 64, 64)),
    "wide_resnet50_2": (_tv("wide_resnet50_2"), "default", (1, 3, 64, 64)),
    "mobilenet_v3_small": (_tv("mobilenet_v3_small"), "default", (1, 3, 64, 64)),
    "efficientnet_b0": (_tv("efficientnet_b0"), "default", (1, 3, 64, 64)),
    "regnet_y_400mf": (_tv("regnet_y_400mf"), "per_channel", (1, 3, 64, 64)),
    "shufflenet_v2_x0_5": (_tv("shufflenet_v2_x0_5"), "default", (1, 3, 64, 64)),
    "inception_v3": (_tv("inception_v3", aux_logits=False, init_weights=False), "default", (1, 3, 96, 96)),
    "convnext_tiny": (_tv("convnext_tiny"), "default", (1, 3, 64, 64)),
    "mobilenet_v3_large_per_channel": (_tv("mobilenet_v3_large"), "per_channel", (1, 3, 64, 64)),
    "resnet34": (_tv("resnet34"), "default", (1, 3, 64, 64)),
    "transformer_block": (TransformerBlock, "default", (2, 10, 32)),
    "transformer_block_per_channel": (TransformerBlock, "per_channel", (2, 10, 32)),
    "concat_net": (ConcatNet, "default", (1, 3, 16, 16)),
    "concat_net_per_channel": (ConcatNet, "per_channel", (1, 3, 16, 16)),
}
