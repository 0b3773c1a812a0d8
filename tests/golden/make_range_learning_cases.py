"""Cases and seeded inputs of the range-learning goldens (shared by the generator and the tests; no reference imports)."""
import torch

# name, shape, dtype, bw, use_symmetric, strict, unsigned, per_channel axis (None = per tensor), (mu, sigma), enc range
CASES = [
    ("pt_asym8", (4099,), "fp32", 8, False, False, False, None, (0.3, 1.0), (-1.7, 2.2)),
    ("pt_asym4", (257, 9), "fp32", 4, False, False, False, None, (0.0, 1.0), (-0.9, 1.4)),
    ("pt_asym16", (1031,), "fp32", 16, False, False, False, None, (2.0, 2.0), (-3.0, 7.0)),
    ("pt_sym8", (2050,), "fp32", 8, True, False, False, None, (0.0, 1.0), (-2.0, 2.0)),
    ("pt_sym8_strict", (2050,), "fp32", 8, True, True, False, None, (0.0, 1.0), (-2.0, 2.0)),
    ("pt_usym8", (1500,), "fp32", 8, True, False, True, None, (1.0, 0.5), (0.0, 2.5)),
    ("pt_gate", (515,), "fp32", 8, False, False, False, None, (0.0, 1.0), (0.25, -0.5)),   # gating must repair this
    ("pc_axis0", (16, 3, 3, 3), "fp32", 8, True, False, False, 0, (0.0, 0.2), (-0.5, 0.5)),
    ("pc_axis0_asym", (10, 7, 5), "fp32", 4, False, False, False, 0, (0.1, 0.3), (-0.4, 0.7)),
    ("pc_axis1", (6, 8, 2, 2), "fp32", 8, True, False, False, 1, (0.0, 0.2), (-0.5, 0.5)),
    ("pc_1d", (33,), "fp32", 8, False, False, False, 0, (0.0, 1.0), (-1.0, 1.5)),
    ("pc_1d_sym", (33,), "fp32", 8, True, False, False, 0, (0.0, 1.0), (-1.0, 1.0)),
    ("pt_bf16_8", (2051,), "bf16", 8, False, False, False, None, (0.0, 1.0), (-1.5, 2.0)),
    ("pt_bf16_sym8", (1027,), "bf16", 8, True, False, False, None, (0.0, 1.0), (-2.0, 2.0)),
    ("pt_bf16_16", (1027,), "bf16", 16, False, False, False, None, (0.0, 1.0), (-2.0, 2.5)),
    ("pc_bf16", (12, 5, 3), "bf16", 8, True, False, False, 0, (0.0, 0.2), (-0.5, 0.5)),
    ("pc_bf16_asym4", (12, 20), "bf16", 4, False, False, False, 0, (0.0, 0.2), (-0.3, 0.5)),
]


def make_inputs(idx, shape, dtype, axis, dist, enc):
    g = torch.Generator().manual_seed(1000 + idx)
    x = torch.randn(shape, generator=g) * dist[1] + dist[0]
    grad = torch.randn(shape, generator=g)
    c = 1 if axis is None else shape[axis]
    spread = 1.0 + 0.1 * torch.arange(c, dtype=torch.float32)
    mn = torch.full((c,), enc[0]) * spread
    mx = torch.full((c,), enc[1]) * spread
    flat = x.view(-1)
    # a few awkward values: exact ties on the grid, far out-of-range, zero, negative zero
    if flat.numel() > 64 and enc[1] > enc[0]:
        delta = (enc[1] - enc[0]) / 255.0
        flat[3], flat[4], flat[5], flat[6], flat[7] = 0.5 * delta, 1.5 * delta, -2.5 * delta, 0.0, -0.0
        flat[8], flat[9] = 1e30, -1e30
    dt = torch.float32 if dtype == "fp32" else torch.bfloat16
    return x.to(dt), grad.to(dt), mn.to(dt), mx.to(dt)


# ---- whole-sim cases (tests/golden/make_range_learning_sim_golden.py) -------------------------------------------------
# name: (torchvision architecture, quantsim config file or None, initialisation scheme, input shape)
SIM_CASES = {
    "resnet18_default_tf": ("resnet18", None, "tf", (2, 3, 64, 64)),
    "resnet18_perchannel_tfe": ("resnet18", "default_config_per_channel.json", "tf_enhanced", (2, 3, 64, 64)),
}


def sim_model(arch):
    import torchvision
    torch.manual_seed(0)
    return getattr(torchvision.models, arch)().eval()


def sim_inputs(shape):
    torch.manual_seed(1)
    x = torch.randn(*shape)
    x2 = torch.randn(*shape) * 1.5
    target = torch.randn(shape[0], 1000)
    return x, x2, target
