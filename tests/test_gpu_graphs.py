"""Small-batch path: Engine.forward_graphed (adjacency build + T iterations + readout captured in one CUDA graph per
batch shape) == the eager forward bit for bit, across replays with new features, across two shapes, and after a weight
update (the graph reads the weight buffer in place)."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ignnition_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("one_launch", [True, False])
def test_forward_graphed_matches_eager_and_oracle(one_launch):
    """``one_launch`` False: the per-stage kernels inside the graph (what batches above 8192 rows capture)"""
    from ignnition_b200 import Engine, ModelDescription, synthetic
    from ignnition_b200.generator import sample_to_tensors
    g = load_golden("routenet_nsfnet")
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    eng = Engine(md, device="cuda:0", seed=0)
    if not one_launch:
        eng.small_graph_rows = 0
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


@pytest.mark.parametrize("case", ["routenet_nsfnet", "qsize_nsfnet", "two_entity_16"])
def test_one_launch_loop_matches_per_stage_fp32_bitwise(case):
    """csrc/small_graph.cu: the whole T-iteration message-passing loop of a small graph in ONE persistent launch
    (weights resident in shared memory, grid barrier between stages) gives the SAME BITS as the per-stage fp32 kernels
    (ign_gru_seq / ign_agg_gru_cell) -- states of every entity and predictions -- and agrees with the fp64 oracle.
    RouteNet (ordered + sum), Q-size (interleave over three entities) and a 16-wide two-source ordered model."""
    from ignnition_b200 import Engine, ModelDescription, ops
    if case == "two_entity_16":
        from test_gpu_model import _two_entity_json, _two_entity_sample, tensors_of
        from ignnition_b200.generator import sample_dimensions
        rng = np.random.RandomState(5)
        mj = _two_entity_json({"type": "ordered"})
        samples = [_two_entity_sample(rng, 6, 5, 9), _two_entity_sample(rng, 12, 7, 40), _two_entity_sample(rng, 3, 2, 1)]
        dims = sample_dimensions(samples[0])
        md = ModelDescription(mj, dims)
        tens = [tensors_of(md, s)[0] for s in samples]
    else:
        g = load_golden(case)
        mj, dims = g["model_json"], g["reference_meta"]["dimensions"]
        md = ModelDescription(mj, dims)
        tens = [orc.normalize_inputs(mj, t) for t in g["reference_tensors"]] * 2
    eng = Engine(md, device="cuda:0", seed=3)
    assert eng._small_program_ok()
    graph = eng.prepare(tens)
    assert graph.small and not graph.order and not graph.meta
    state0 = eng.initial_states(graph)
    l0 = eng.gpu_launches
    state = eng.message_passing(graph, state0)
    assert eng.gpu_launches - l0 == 1                       # the whole loop
    pred = eng.readout_forward(state, g=graph)
    ref_eng = Engine(md, device="cuda:0", seed=3)
    ref_eng.small_graph_rows = 0
    ops.set_tensor_cores(False)
    try:
        g2 = ref_eng.prepare(tens)
        assert not g2.small
        state2 = ref_eng.message_passing(g2, ref_eng.initial_states(g2))
        pred2 = ref_eng.readout_forward(state2, g=g2)
    finally:
        ops.set_tensor_cores(True)
    for e in eng.entities:
        assert torch.equal(state[e], state2[e]), e
    o64 = orc.Oracle(mj, dims, dtype=np.float64)
    w = eng.get_weights()
    want = [o64.forward(t, w, return_states=True) for t in tens]
    # fp32 arithmetic against fp64: 1e-5 everywhere except the Q-size node states, where fp32 ITSELF sits at
    # 1.0e-5 .. 1.6e-5 of the fp64 result (profiles/r2_parity.md: a NumPy fp32 run of the oracle deviates as much)
    tol = 2e-5 if case == "qsize_nsfnet" else 1e-5
    for e in eng.entities:
        ref = np.concatenate([s[e] for _, s in want])
        assert float(np.abs(state[e].cpu().numpy() - ref).max() / np.abs(ref).max()) < tol, e
    ref = np.concatenate([p.reshape(-1) for p, _ in want])
    assert float(np.abs(pred.cpu().numpy().reshape(-1) - ref).max() / np.abs(ref).max()) < tol
    # through the captured graph as well (the barrier counter is reset by a memset node at every replay)
    batch = eng.assemble(tens)
    for _ in range(3):
        assert torch.equal(eng.forward_graphed(batch), pred)
    # a training graph keeps the per-stage kernels (the backward needs the walk's tables)
    assert not eng.prepare(tens, training=True).small


def test_train_step_graphed_matches_eager_steps():
    """Trainer.train_step_graphed (adjacency build + forward + loss + backward as one captured graph per batch shape,
    optimiser update outside) walks the same weights as Trainer.train_step over a stream of batches with new
    topologies and features (the weight-gradient flush uses fp32 atomics: not bit for bit; the per-step losses agree to 1e-5)"""
    from ignnition_b200 import Engine, ModelDescription, synthetic
    from ignnition_b200.generator import sample_to_tensors
    from ignnition_b200.train import Trainer
    g = load_golden("routenet_nsfnet")
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    feats = [f.name for f in md.get_all_features()]

    def sample(seed):
        t, y = sample_to_tensors(synthetic.routenet_sample("nsfnet", seed, seed), feats, "delay",
                                 md.get_adjecency_info(), [], [], True)
        return orc.normalize_inputs(g["model_json"], t), np.log(np.asarray(y, np.float32))

    engines = [Engine(md, device="cuda:0", seed=1) for _ in range(2)]
    trainers = [Trainer(e) for e in engines]
    for step in range(4):
        both = [sample(100 + 3 * step + k) for k in range(3)]
        tens, labels = [b[0] for b in both], [b[1] for b in both]
        batch = engines[0].assemble(tens, labels)
        trainers[0].train_step(engines[0].prepare(batch, training=True))
        trainers[1].train_step_graphed(engines[1].assemble(tens, labels))
        a, b = trainers[0].losses(), trainers[1].losses()
        assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(a["loss"])
    w0, w1 = engines[0].weights.cpu().numpy(), engines[1].weights.cpu().numpy()
    # (same kernels in both; the weight-gradient flushes use atomics, and Adam turns a 1e-9 difference of a near-zero
    # gradient into up to lr * 1e-2 of a weight per step)
    assert float(np.abs(w0 - w1).max() / np.abs(w0).max()) < 1e-4
    assert 1 <= len(trainers[1]._graphs) <= 4 and trainers[1].step == 4
