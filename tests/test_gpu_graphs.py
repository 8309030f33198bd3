"""Small-batch path: Engine.forward_graphed (adjacency build + T iterations + readout captured in one CUDA graph per
batch shape) == the eager forward bit for bit, across replays with new features, across two shapes, and after a weight
update (the graph reads the weight buffer in place)."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ignnition_oracle as orc

pytestmark = pytest.mark.gpu


def test_forward_graphed_matches_eager_and_oracle():
    from ignnition_b200 import Engine, ModelDescription, synthetic
    from ignnition_b200.generator import sample_to_tensors
    g = load_golden("routenet_nsfnet")
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    eng = Engine(md, device="cuda:0", seed=0)
    o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
    feats = [f.name for f in md.get_all_features()]

    def tensors(seed):
        t, _ = sample_to_tensors(synthetic.routenet_sample("nsfnet", seed, seed), feats, "delay", md.get_adjecency_info(),
                                 [], [], True)
        return orc.normalize_inputs(g["model_json"], t)

    for n_samples in (3, 5, 3):                      # the third round replays the graph captured by the first
        for rep in range(2):
            tens = [tensors(10 * n_samples + rep * 7 + k) for k in range(n_samples)]
            batch = eng.assemble(tens)
            got = eng.forward_graphed(batch).clone()
            want = eng.forward(eng.prepare(tens))
            assert torch.equal(got, want)
    n_graphs = len(eng._graphs)              # one per batch shape (row counts AND longest sequences)
    assert 2 <= n_graphs <= 6
    feats2 = eng.assemble(tens)                # the same shape again: no new capture
    assert torch.equal(eng.forward_graphed(feats2), want) and len(eng._graphs) == n_graphs
    w = eng.get_weights()
    ref = np.concatenate([o64.forward(t, w).reshape(-1) for t in tens])
    assert float(np.abs(got.cpu().numpy().reshape(-1) - ref).max() / np.abs(ref).max()) < 1e-5
    # new weights are seen by the captured graph
    eng.reset_parameters(seed=5)
    got2 = eng.forward_graphed(batch).clone()
    assert torch.equal(got2, eng.forward(eng.prepare(tens))) and not torch.equal(got2, got)
