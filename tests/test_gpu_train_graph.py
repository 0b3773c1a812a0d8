"""QuantizationSimModel.capture_train_step: a QAT step replayed from a CUDA graph must train like the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup():
    from aimet_b200.quantsim import QuantizationSimModel
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.BatchNorm2d(16), torch.nn.ReLU(),
                                torch.nn.Conv2d(16, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1),
                                torch.nn.Flatten(), torch.nn.Linear(8, 4)).cuda()
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(8, 3, 16, 16, generator=g).cuda() for _ in range(4)]
    ys = [torch.randint(0, 4, (8,), generator=g).cuda() for _ in range(4)]
    sim = QuantizationSimModel(model, dummy_input=xs[0], quant_scheme="tf_enhanced", in_place=True)
    sim.compute_encodings(lambda m, _: m(xs[0]), None)
    sim.model.train()
    opt = torch.optim.SGD(sim.model.parameters(), lr=0.05, momentum=0.9)
    return sim, opt, xs, ys


def _loss(out, y):
    return torch.nn.functional.cross_entropy(out, y)


def test_graphed_qat_step_trains_like_the_eager_step():
    sim_e, opt_e, xs, ys = _setup()
    losses_e = []
    for x, y in zip(xs, ys):
        opt_e.zero_grad(set_to_none=True)
        loss = _loss(sim_e.model(x), y)
        loss.backward()
        opt_e.step()
        losses_e.append(float(loss))
    sim_g, opt_g, xs, ys = _setup()
    before = {n: p.detach().clone() for n, p in sim_g.model.named_parameters()}
    step = sim_g.capture_train_step(_loss, opt_g, (xs[0],), ys[0])
    for n, p in sim_g.model.named_parameters():
        assert torch.equal(p.detach(), before[n]), n           # capturing (with its warm-up steps) left the model untouched
    losses_g = [float(step(x, target=y)) for x, y in zip(xs, ys)]
    assert losses_g == pytest.approx(losses_e, rel=1e-5)
    assert losses_g[-1] != losses_g[0]
    for (n, a), (_, b) in zip(sim_e.model.named_parameters(), sim_g.model.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), n
    for (n, a), (_, b) in zip(sim_e.model.named_buffers(), sim_g.model.named_buffers()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=1e-6), n


def test_graphed_ddp_qat_step_trains_like_the_eager_ddp_step():
    """capture_train_step(model=<DDP wrapper>): the reducer's all-reduces are captured with the backward. Runs in a helper
    process (NCCL needs its environment set before the process group exists)."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, os.path.join(here, "ddp_graph_driver.py"), "29541"], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=600, cwd=os.path.dirname(here))
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line, res.stdout[-3000:]
    out = json.loads(line[-1][len("RESULT "):])
    assert out["graphed"] == pytest.approx(out["eager"], rel=1e-5)
    assert out["graphed"][-1] != out["graphed"][0]
    assert out["parameters_close"]
