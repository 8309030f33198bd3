"""GPU parity of the train step (model_fn): gradients vs torch.autograd on the differentiable CPU
oracle, Adam update vs the restated Keras formula."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ignnition_oracle as orc
from oracle.torch_port import TorchOracle

pytestmark = pytest.mark.gpu
GRAD_RTOL = 2e-4      # fp32 backward through T=8 x BPTT vs an fp64 autograd reference (normwise)
# selu' jumps from 1.0507 to 1.7581 at 0: ONE readout pre-activation (of ~1e5) that sits within the
# forward tolerance of zero and lands on the other side flips that row's whole contribution to the
# gradients of the layers below it (measured: 6.8e-4 on the 2nd readout layer from one flip with the
# 3xTF32 forward at 2.7e-6; the fp32 twin kernels at 6e-7 do not flip it).  TF's own fp32 kernels are
# exposed to the same jump, so the layers under a selu get this bound instead.
GRAD_RTOL_SELU_KINK = 1.5e-3


def rel_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def setup(case, n_samples=None):
    from ignnition_b200 import Engine, ModelDescription
    from ignnition_b200.train import Trainer
    g = load_golden(case)
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
    w = {k: v.astype(np.float32) for k, v in o64.init_weights(1234).items()}
    eng = Engine(md, device="cuda:0")
    eng.set_weights(w)
    tens = [orc.normalize_inputs(g["model_json"], t) for t in g["reference_tensors"]][:n_samples]
    out_name, out_norm, _ = md.get_output_info()
    labels = [np.asarray(orc.EXAMPLE_NORMALIZATIONS[out_norm](np.asarray(y, np.float32), out_name), np.float32)
              for y in g["reference_labels"]][:n_samples]
    return g, dims, eng, Trainer(eng), w, tens, labels


@pytest.mark.parametrize("case", ["routenet_nsfnet", "qsize_hand", "qsize_nsfnet"])
def test_gradients_match_autograd(case):
    g, dims, eng, tr, w, tens, labels = setup(case)
    graph = eng.prepare(tens, labels=labels, training=True)
    pred, n_local = tr.loss_and_grads(graph)
    for name, lam in eng._reg.items():
        from ignnition_b200 import ops
        ops.l2_reg(eng.param(name), lam, tr.g(name), tr.scalars[1:2])
    mse, reg, p_ref, grads = TorchOracle(g["model_json"], dims).loss_and_grads(tens, labels, w)
    sc = tr.scalars.cpu().numpy()
    assert rel_err(pred.cpu().numpy().reshape(-1), p_ref) < 1e-5
    assert abs(sc[0] / n_local - mse) <= 1e-5 * abs(mse)
    assert abs(sc[1] - reg) <= 1e-5 * abs(reg)
    got = tr.grads.cpu().numpy()
    for name, (off, shape) in eng.param_table.items():
        gn = got[off:off + int(np.prod(shape))].reshape(shape)
        tol = GRAD_RTOL_SELU_KINK if name.startswith("readout_model") else GRAD_RTOL
        assert rel_err(gn, grads[name]) < tol, name


def test_adam_step_matches_keras_formula():
    g, dims, eng, tr, w, tens, labels = setup("routenet_nsfnet", 1)
    graph = eng.prepare(tens, labels=labels, training=True)
    w0 = eng.weights.clone()
    tr.train_step(graph)
    grads = tr.grads.cpu().numpy().astype(np.float64)
    lr = orc.exponential_decay(0, 1e-3, 80000, 0.6)
    want, _, _ = orc.adam_step(w0.cpu().numpy().astype(np.float64), grads, 0.0, 0.0, 1, lr)
    assert np.abs(eng.weights.cpu().numpy() - want).max() < 1e-7
    l1 = tr.losses()
    for _ in range(20):
        tr.train_step(graph)
    l2 = tr.losses()
    assert l2["total_loss"] < l1["total_loss"]            # the step actually descends
    assert tr.step == 21


def test_training_unsupported_shapes_fail_loudly():
    from ignnition_b200 import ops
    z = torch.zeros(8, 64, device="cuda")
    w = torch.zeros(64, 192, device="cuda")
    b = torch.zeros(2, 192, device="cuda")
    with pytest.raises(RuntimeError, match="IGNNITION.*backward pass is built for"):
        ops.gru_cell_bwd(z, z, w, w, b, z, z.clone(), z.clone(), w.clone(), w.clone(), b.clone())
