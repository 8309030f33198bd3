"""CPU tests: host logic and the oracle against the golden vectors produced by the reference's own
generator / parser (``oracle/make_golden.py``), and the C-ABI export check."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from ignnition_b200 import synthetic
from ignnition_b200.batching import AdjacencySpec, SequenceSpec, assemble, assemble_tiled, position_table
from ignnition_b200.generator import interleave_indices, make_indices, sample_dimensions, sample_to_tensors
from ignnition_b200.model_description import ModelDescription, ModelDescriptionError
from oracle import ignnition_oracle as orc

CASES = ["routenet_nsfnet", "qsize_hand", "qsize_nsfnet", "routenet_geant2", "routenet_synth50"]


def _samples(g):
    if g["samples"] is not None:
        return g["samples"]
    shape, ts, fs = g["sample_recipe"]
    return [synthetic.routenet_sample(shape, ts, fs)]


def _tensors(g, sample, md):
    feats = [f.name for f in md.get_all_features()]
    out, _, _ = md.get_output_info()
    return sample_to_tensors(sample, feats, out, md.get_adjecency_info(), md.get_interleave_tensors(),
                             md.get_additional_input_names(), True)


# ------------------------------------------------------------------ parser vs the reference parser
@pytest.mark.parametrize("case", CASES)
def test_parser_matches_reference(case):
    g = load_golden(case)
    meta = g["reference_meta"]
    md = ModelDescription(g["model_json"], meta["dimensions"])
    assert md.get_adjecency_info() == meta["adjacency_info"]
    assert md.get_interleave_tensors() == meta["interleave_tensors"]
    assert md.get_interleave_sources() == meta["interleave_sources"]
    assert [[f.name, f.size, f.normalization] for f in md.get_all_features()] == meta["features"]
    assert list(md.get_output_info()) == meta["output_info"]
    assert md.get_mp_iterations() == meta["mp_iterations"]
    assert md.get_input_dimensions() == meta["input_dimensions"]
    assert md.get_additional_input_names() == meta["additional_input"]
    assert md.get_loss() == meta["loss"]
    assert md.get_optimizer() == meta["optimizer"]
    stages = [[name, [[mp.destination_entity, [s.name for s in mp.source_entities], mp.aggregation.type,
                       mp.update.type] for mp in mps]] for name, mps in md.get_mp_instances()]
    assert stages == meta["stages"]


def test_parser_rejects_bad_models():
    g = load_golden("qsize_hand")
    bad = dict(g["model_json"])
    bad = {k: v for k, v in bad.items() if k != "readout"}
    with pytest.raises(ModelDescriptionError, match="IGNNITION"):
        ModelDescription(bad)
    import copy
    bad = copy.deepcopy(g["model_json"])
    bad["message_passing"]["stages"][0]["stage_mp"][0]["destination_entity"] = "nope"
    with pytest.raises(ModelDescriptionError, match="destination entity nope"):
        ModelDescription(bad)
    bad = copy.deepcopy(g["model_json"])
    bad["message_passing"]["stages"][0]["stage_mp"][0]["update"]["nn_name"] = "missing_nn"
    with pytest.raises(ModelDescriptionError, match="missing_nn"):
        ModelDescription(bad)


# ------------------------------------------------------------------ generator vs the reference generator
@pytest.mark.parametrize("case", CASES)
def test_generator_bit_exact(case):
    g = load_golden(case)
    md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
    for sample, ref, ref_y in zip(_samples(g), g["reference_tensors"], g["reference_labels"]):
        assert sample_dimensions(sample) == g["reference_meta"]["dimensions"]
        got, y = _tensors(g, sample, md)
        assert set(got) == set(ref)
        for k, v in ref.items():
            if k.startswith(("src_", "dst_", "seq_", "indices_", "num_")):
                assert np.array_equal(np.asarray(got[k], dtype=np.int64), np.asarray(v, dtype=np.int64)), k
            else:
                assert np.array_equal(np.asarray(got[k], dtype=np.float64), np.asarray(v, dtype=np.float64)), k
        assert np.allclose(y, ref_y, rtol=0, atol=0)


def test_make_indices_order_of_appearance():
    cnt, idx = make_indices({"a": "x", "b": "y", "c": "x", "d": "y", "e": "x"})
    assert cnt == {"x": 3, "y": 2}
    assert idx == {"a": 0, "b": 0, "c": 1, "d": 1, "e": 2}


def test_interleave_hand_example():
    # pattern [node, link], max lens 3/3 -> node [0,2,4], link [1,3,5]
    assert interleave_indices(["node", "link"], {"node": 3, "link": 3}) == {"node": [0, 2, 4], "link": [1, 3, 5]}
    assert interleave_indices(["node", "link"], {"node": 4, "link": 3}) == {"node": [0, 2, 4, 6], "link": [1, 3, 5]}


def test_generator_errors():
    g = load_golden("qsize_hand")
    md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
    s = dict(g["samples"][0])
    del s["traffic"]
    with pytest.raises(Exception, match='feature named "traffic"'):
        _tensors(g, s, md)
    s = dict(g["samples"][0])
    s["entities"] = dict(s["entities"], p0="link")
    with pytest.raises(Exception, match="adjecency list"):
        _tensors(g, s, md)


# ------------------------------------------------------------------ oracle: integer side
@pytest.mark.parametrize("case", CASES)
def test_oracle_csr_roundtrip(case):
    g = load_golden(case)
    for ref in g["reference_tensors"]:
        for adj, src_e, dst_e, _ in g["reference_meta"]["adjacency_info"]:
            src, dst = np.array(ref["src_" + adj]), np.array(ref["dst_" + adj])
            seq = np.array(ref["seq_%s_%s" % (src_e, dst_e)])
            n = ref["num_" + dst_e]
            rowptr, col, perm = orc.csr_from_edges(src, dst, seq, n)
            assert rowptr[-1] == len(src) and np.all(col >= 0)
            # CSR -> (dst, seq, src) round trip
            d2 = np.repeat(np.arange(n), np.diff(rowptr))
            s2 = np.arange(len(src)) - rowptr[d2]
            assert np.array_equal(d2, dst[perm]) and np.array_equal(s2, seq[perm]) and np.array_equal(col, src[perm])
            # stable sort by destination gives the same CSR (seq is the rank in input order)
            r2, c2, p2 = orc.stable_sort_csr(src, dst, n)
            assert np.array_equal(r2, rowptr) and np.array_equal(c2, col) and np.array_equal(p2, perm)


def test_oracle_segment_sum_equals_padded_reduce_sum():
    rng = np.random.RandomState(0)
    n_dst, n_src, F = 17, 11, 8
    lens = rng.randint(0, 6, n_dst)
    dst = np.repeat(np.arange(n_dst), lens)
    seq = np.concatenate([np.arange(l) for l in lens])
    p = rng.permutation(len(dst))
    dst, seq = dst[p], seq[p]
    src = rng.randint(0, n_src, len(dst))
    states = rng.randn(n_src, F).astype(np.float32)
    padded = np.zeros((n_dst, lens.max(), F), np.float32)
    padded[dst, seq] = states[src]
    rowptr, col, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    seg = np.stack([states[col[rowptr[d]:rowptr[d + 1]]].sum(axis=0) if lens[d] else np.zeros(F, np.float32)
                    for d in range(n_dst)])
    assert np.allclose(seg, padded.sum(axis=1), rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ oracle: float side cross-checks
def test_oracle_gru_matches_torch_grucell():
    rng = np.random.RandomState(1)
    u, fi, n = 32, 32, 50
    K = rng.uniform(-0.3, 0.3, (fi, 3 * u)).astype(np.float32)
    R = rng.uniform(-0.3, 0.3, (u, 3 * u)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (2, 3 * u)).astype(np.float32)
    x = rng.randn(n, fi).astype(np.float32)
    h = rng.randn(n, u).astype(np.float32)
    got = orc.gru_cell(x, h, K, R, b)
    cell = torch.nn.GRUCell(fi, u)
    perm = np.concatenate([np.arange(u, 2 * u), np.arange(0, u), np.arange(2 * u, 3 * u)])   # z|r|h -> r|z|n
    with torch.no_grad():
        cell.weight_ih.copy_(torch.from_numpy(K[:, perm].T.copy()))
        cell.weight_hh.copy_(torch.from_numpy(R[:, perm].T.copy()))
        cell.bias_ih.copy_(torch.from_numpy(b[0, perm].copy()))
        cell.bias_hh.copy_(torch.from_numpy(b[1, perm].copy()))
        ref = cell(torch.from_numpy(x), torch.from_numpy(h)).numpy()
    assert np.abs(got - ref).max() < 2e-6


def test_oracle_selu_matches_torch():
    x = np.linspace(-6, 6, 101).astype(np.float32)
    ref = torch.nn.functional.selu(torch.from_numpy(x)).numpy()
    assert np.abs(orc.activation("selu", x) - ref).max() < 1e-6


def test_oracle_masked_rnn_is_per_destination_walk():
    rng = np.random.RandomState(2)
    n, L, F = 9, 5, 8
    lens = rng.randint(1, L + 1, n)
    x = rng.randn(n, L, F).astype(np.float32)
    h0 = rng.randn(n, F).astype(np.float32)
    K = rng.uniform(-0.4, 0.4, (F, 3 * F)).astype(np.float32)
    R = rng.uniform(-0.4, 0.4, (F, 3 * F)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (2, 3 * F)).astype(np.float32)
    cell = lambda a, h: orc.gru_cell(a, h, K, R, b)
    got = orc.masked_rnn_last(cell, x, h0, lens)
    for i in range(n):
        h = h0[i:i + 1]
        for t in range(lens[i]):
            h = cell(x[i:i + 1, t], h)
        assert np.allclose(got[i], h[0], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        orc.masked_rnn_last(cell, x, h0, np.zeros(n, dtype=np.int64))


@pytest.mark.parametrize("case", ["routenet_nsfnet", "qsize_hand", "qsize_nsfnet"])
def test_oracle_float_golden(case):
    g = load_golden(case)
    dims = g["reference_meta"]["dimensions"]
    for ref, fl in zip(g["reference_tensors"], g["oracle_float"]):
        o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
        w = {k: v.astype(np.float32) for k, v in o64.init_weights(fl["weight_seed"]).items()}
        tens = orc.normalize_inputs(g["model_json"], ref)
        p64, st = o64.forward(tens, w, return_states=True)
        assert np.allclose(p64.reshape(-1), fl["predictions_fp64"], rtol=1e-12, atol=1e-12)
        for k, v in fl["state_checksums_fp64"].items():
            assert abs(st[k].sum() - v) <= 1e-9 * max(1.0, abs(v))
        # fp32 run of the same program stays within the north-star tolerance of the fp64 shadow
        p32 = orc.Oracle(g["model_json"], dims, dtype=np.float32).forward(tens, w)
        err = np.abs(p32.reshape(-1) - p64.reshape(-1)).max() / np.abs(p64).max()
        assert err < 1e-5, err


def test_oracle_weight_count_routenet():
    g = load_golden("routenet_nsfnet")
    o = orc.Oracle(g["model_json"], g["reference_meta"]["dimensions"])
    total = sum(int(np.prod(s)) for s in o.weight_shapes().values())
    assert total == 87169           # 2 x 6336 GRU + 8448 + 65792 + 257


def test_exponential_decay_and_adam():
    assert orc.exponential_decay(80000, 1e-3, 80000, 0.6) == pytest.approx(6e-4)
    assert orc.exponential_decay(81999, 1e-3, 82000, 0.8, staircase="True") == pytest.approx(1e-3)
    w, g = np.ones(3), np.full(3, 0.5)
    w1, m, v = orc.adam_step(w, g, np.zeros(3), np.zeros(3), 1, 1e-3)
    assert np.allclose(w1, 1 - 1e-3 * 0.5 / (0.5 + 1e-7 * np.sqrt(1 - 0.999) / 1.0) , rtol=1e-5)


# ------------------------------------------------------------------ batching
def _specs(md):
    ents = [e.name for e in md.get_entities()]
    feats = [(f.name, e.name, f.size) for e in md.get_entities() for f in e.features]
    adjs = [AdjacencySpec(a[0], a[1], a[2], a[3] == "True") for a in md.get_adjecency_info()]
    by = {a.name: a for a in adjs}
    seqs = []
    for si, (_, mps) in enumerate(md.get_mp_instances()):
        for mi, mp in enumerate(mps):
            if mp.aggregation.type == "interleave" or (mp.aggregation.type == "ordered" and len(mp.source_entities) > 1):
                seqs.append(SequenceSpec("s%d_m%d" % (si, mi), mp.destination_entity,
                                         [by[s.adj_vector] for s in mp.source_entities],
                                         mp.aggregation.type == "interleave"))
    return ents, feats, adjs, seqs


def test_batch_is_block_diagonal():
    g = load_golden("routenet_nsfnet")
    md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
    ents, feats, adjs, seqs = _specs(md)
    samples = g["reference_tensors"]
    b = assemble(samples, ents, feats, adjs, seqs, g["reference_labels"])
    assert b.num["path"] == sum(s["num_path"] for s in samples)
    off = 0
    for s in samples:
        n = len(s["src_adj_links_paths"])
        assert np.array_equal(b.arrays["src_adj_links_paths"][off:off + n] - b.offsets["link"][samples.index(s)],
                              s["src_adj_links_paths"])
        off += n
    assert b.arrays["labels"].size == b.num["path"]
    # tiled assembly == assembling copies
    t = assemble_tiled(samples[0], 3, ents, feats, adjs, seqs,
                       {"traffic": lambda r, n: np.zeros(n), "link_capacity": lambda r, n: np.zeros(n)})
    c = assemble([samples[0]] * 3, ents, feats, adjs, seqs)
    for k in ("src_adj_links_paths", "dst_adj_links_paths", "seq_adj_links_paths", "src_adj_paths_links",
              "dst_adj_paths_links", "sample_of_path"):
        assert np.array_equal(t.arrays[k], c.arrays[k]), k


def test_position_table_interleave():
    g = load_golden("qsize_hand")
    md = ModelDescription(g["model_json"], g["reference_meta"]["dimensions"])
    _, _, _, seqs = _specs(md)
    ps, pc = position_table(g["reference_tensors"][0], seqs[0])
    # sources in model order: link (k=0), node (k=1); pattern [node, link] -> n0 l0 n1 l1 n2 l2
    assert ps.tolist() == [1, 0, 1, 0, 1, 0]
    assert pc.tolist() == [0, 0, 1, 1, 2, 2]


# ------------------------------------------------------------------ the C-ABI library
def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ignnition_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ign_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    from ignnition_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared
    path = _lib.LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ign_version() >= 100
    # host-side argument validation works without a GPU: invalid calls return <0 and set the message
    lib.ign_segment_reduce.restype = ctypes.c_int
    rc = lib.ign_segment_reduce(99, None, None, None, 32, ctypes.c_int64(1), None, None)
    assert rc < 0
    buf = ctypes.create_string_buffer(256)
    lib.ign_last_error(buf, ctypes.c_size_t(256))
    assert buf.value.decode().startswith("IGNNITION:")


# ------------------------------------------------------------------ the differentiable oracle twin
@pytest.mark.parametrize("case", ["routenet_nsfnet", "qsize_hand"])
def test_torch_oracle_matches_numpy_oracle(case):
    from oracle.torch_port import TorchOracle
    g = load_golden(case)
    dims = g["reference_meta"]["dimensions"]
    o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
    w = {k: v.astype(np.float32) for k, v in o64.init_weights(1234).items()}
    tens = orc.normalize_inputs(g["model_json"], g["reference_tensors"][0])
    to = TorchOracle(g["model_json"], dims)
    p_t = to.forward(tens, to.params(w)).detach().numpy()
    p_n = o64.forward(tens, w)
    assert np.abs(p_t - p_n).max() < 1e-12


def test_learning_rate_schedules():
    from ignnition_b200.train import LearningRate
    lr = LearningRate({"type": "Adam", "schedule": {"type": "ExponentialDecay", "initial_learning_rate": 0.001,
                                                   "decay_steps": 82000, "decay_rate": 0.8, "staircase": "True"}})
    assert lr(0) == pytest.approx(1e-3) and lr(81999) == pytest.approx(1e-3) and lr(82000) == pytest.approx(8e-4)
    lr = LearningRate({"type": "Adam", "schedule": {"type": "ExponentialDecay", "initial_learning_rate": 0.001,
                                                   "decay_steps": 80000, "decay_rate": 0.6}})
    assert lr(40000) == pytest.approx(orc.exponential_decay(40000, 1e-3, 80000, 0.6))
    assert LearningRate({"type": "Adam", "learning_rate": 0.01})(5) == 0.01


def test_cabi_rejects_oversized_and_bad_arguments_on_the_host():
    """argument validation happens before any launch, so it is testable without a GPU"""
    import ctypes as C
    from ignnition_b200 import _lib
    lib = _lib.load()
    one = C.c_void_p(16)          # never dereferenced: the size checks come first
    rc = lib.ign_csr_build(one, one, None, 1 << 31, 10, 0, one, one, None, None, one, 1 << 40, None)
    assert rc == -2 and "int32" in _lib.last_error()                    # IGN_ERR_UNSUPPORTED: maximum size
    rc = lib.ign_csr_build(one, one, None, 100, 10, 0, one, one, None, None, one, 16, None)
    assert rc == -3 and "workspace" in _lib.last_error()                # IGN_ERR_WORKSPACE
    rc = lib.ign_gru_seq(one, one, None, 9, one, 32, one, 10, 32, one, one, one, one, None, None, None)
    assert rc == -1                                                      # too many sources
    rc = lib.ign_gru_cell(one, one, 10, 48, 48, one, one, one, one, None, 0, None)
    assert rc == -2 and "not built" in _lib.last_error()
    assert lib.ign_csr_build_ws_bytes(200_000_000, 10_000_000) > 3_200_000_000   # 4 x E x 4 B of sort buffers
    assert lib.ign_dense_ws_bytes(256, 256) == 8 * 2 * 256 * 128 and lib.ign_dense_ws_bytes(65, 20) == 0
    assert lib.ign_dense_bwd_ws_bytes(32, 256) == lib.ign_dense_ws_bytes(256, 32) and lib.ign_dense_bwd_ws_bytes(256, 1) == 0
    rc = lib.ign_attention_aggregate(one, one, None, one, 30, one, one, one, 1, 10, 20, 4, one, one, 1 << 20, None)
    assert rc == -2 and "multiple of 4" in _lib.last_error()            # message width
    rc = lib.ign_attention_aggregate(one, one, None, one, 32, one, one, one, 1, 10, 20, 4, one, one, 16, None)
    assert rc == -1 and "workspace" in _lib.last_error()
    assert lib.ign_attention_ws_bytes(1000, 4, 8) >= 4 * 8 * 16 + 1000 * 4
    assert lib.ign_attention_combine(9, one, one, one, 4, 2, one, one, one, None) == -2       # at most 8 sources
    assert lib.ign_mul(-1, one, one, one, None) == -1 and lib.ign_mul(0, None, None, None, None) == 0
    assert lib.ign_conv_finish(one, one, one, 0, 10, 1, one, None) == -1
    assert lib.ign_partner_index(None, None, None, 5, None, None) == -1 and lib.ign_partner_index(None, None, None, 0, None, None) == 0
    with pytest.raises(RuntimeError, match="IGNNITION"):
        _lib.check(-1, "x")


class _SpecEngine:
    """what NativeIngest reads from an Engine (no GPU needed to build it)"""

    def __init__(self, md):
        from ignnition_b200.batching import AdjacencySpec
        self.entities = [e.name for e in md.get_entities()]
        self.features = [(f.name, e.name, f.size) for e in md.get_entities() for f in e.features]
        from ignnition_b200.batching import SequenceSpec
        self.model = md
        self.adjacencies, self.sequences = [], []
        by_name = {}
        for si, (_, mps) in enumerate(md.get_mp_instances()):
            for mi, mp in enumerate(mps):
                for src in mp.source_entities:
                    if src.adj_vector not in by_name:
                        uses = any(i == "edge_params" for op in src.message_formation for i in getattr(op, "input", []) or [])
                        by_name[src.adj_vector] = AdjacencySpec(src.adj_vector, src.name, mp.destination_entity, uses)
                        self.adjacencies.append(by_name[src.adj_vector])
                agg = mp.aggregation.type
                if agg == "interleave" or (agg in ("ordered", "concat") and len(mp.source_entities) > 1):
                    self.sequences.append(SequenceSpec("s%d_m%d" % (si, mi), mp.destination_entity,
                                                       [by_name[s.adj_vector] for s in mp.source_entities],
                                                       agg == "interleave"))


@pytest.mark.parametrize("case", ["routenet_nsfnet", "routenet_geant2"])
def test_native_ingest_equals_python_generator(case):
    """the C++ ingest of data.json text gives, array for array, what the Python mirror of
    the reference generator + block-diagonal assembly gives (integer arrays bit-exact, floats bit-exact)."""
    import json
    from ignnition_b200 import synthetic
    from ignnition_b200.batching import assemble
    from ignnition_b200.generator import sample_dimensions, sample_to_tensors
    from ignnition_b200.ingest import NativeIngest
    from ignnition_b200.model_description import ModelDescription
    g = load_golden(case)
    topo = "geant2" if "geant2" in case else "nsfnet"
    samples = [synthetic.routenet_sample(topo, s, s) for s in range(5)] + list((g.get("samples") or [])[:1])
    md = ModelDescription(g["model_json"], sample_dimensions(samples[0]))
    eng = _SpecEngine(md)
    out_name = md.get_output_info()[0]
    feats = [f[0] for f in eng.features]
    adj = md.get_adjecency_info()
    pairs = [sample_to_tensors(s, feats, out_name, adj, [], [], True) for s in samples]
    want = assemble([p[0] for p in pairs], eng.entities, eng.features, eng.adjacencies, (), [p[1] for p in pairs])
    ing = NativeIngest(eng, label_name=out_name)
    text = json.dumps(samples)
    assert ing.parse(text) == len(samples)
    got = ing.batch()
    assert got.n_samples == want.n_samples and got.num == want.num and got.n_edges == want.n_edges
    assert got.max_seq == want.max_seq
    assert set(got.arrays) == set(want.arrays)
    for k in want.arrays:
        assert got.arrays[k].dtype == want.arrays[k].dtype, k
        assert np.array_equal(got.arrays[k], want.arrays[k]), k
    for e in eng.entities:
        assert np.array_equal(got.offsets[e], want.offsets[e])
    # one object at a time appends to the same batch; reset starts over
    ing.reset()
    for s in samples:
        assert ing.parse(json.dumps(s, indent=1)) == 1          # whitespace-tolerant
    again = ing.batch()
    for k in want.arrays:
        assert np.array_equal(again.arrays[k], want.arrays[k]), k


def test_native_ingest_edge_params_and_errors():
    import json
    from ignnition_b200.batching import AdjacencySpec
    from ignnition_b200.ingest import NativeIngest

    class E:
        entities = ["node"]
        features = [("x", "node", 2)]
        adjacencies = [AdjacencySpec("adj", "node", "node", True)]
        sequences = []

    s = {"entities": {"b": "node", "a\u00e9": "node", "c": "node"}, "x": [[1, 2.5], [3e-1, -4], [0.1, 7]],
         "adj": {"c": [["b", [1.9, -2.9]], ["a\u00e9", [3, 4]]], "b": []}, "y": 3.25}
    ing = NativeIngest(E(), label_name="y")
    assert ing.parse(json.dumps(s)) == 1 and ing.parse(json.dumps([s, s])) == 2
    b = ing.batch()
    assert b.n_samples == 3 and b.num["node"] == 9
    assert np.array_equal(b.arrays["src_adj"], [0, 1, 3, 4, 6, 7]) and np.array_equal(b.arrays["dst_adj"], [2, 2, 5, 5, 8, 8])
    assert np.array_equal(b.arrays["seq_adj"], [0, 1] * 3)
    assert np.array_equal(b.arrays["params_adj"][:2], [[1.0, -2.0], [3.0, 4.0]])       # truncation, quirk 11
    assert np.array_equal(b.arrays["feat_x"][:6], np.asarray([1, 2.5, 3e-1, -4, 0.1, 7], np.float32))
    assert np.array_equal(b.arrays["labels"], [3.25] * 3)
    bad = dict(s); bad["adj"] = {"zzz": ["b"]}
    ing.reset()
    with pytest.raises(RuntimeError, match="not in the entities"):
        ing.parse(json.dumps(bad))
    ing.reset()
    with pytest.raises(RuntimeError, match="was not found although being expected"):
        ing.parse(json.dumps({"entities": {"a": "node"}, "x": [[1, 2]]}))
    with pytest.raises(RuntimeError, match="malformed"):
        ing.parse('{"entities": {"a": "node"}, "x": [[1, 2]], "adj": {"a": ["a",]')


def test_native_ingest_reads_dataset_files_in_parallel(tmp_path):
    import json
    from ignnition_b200 import synthetic
    from ignnition_b200.batching import assemble
    from ignnition_b200.generator import read_dataset, sample_dimensions, sample_to_tensors
    from ignnition_b200.ingest import NativeIngest
    from ignnition_b200.model_description import ModelDescription
    g = load_golden("routenet_nsfnet")
    samples = [synthetic.routenet_sample("nsfnet", k % 3, k) for k in range(10)]
    synthetic.write_dataset(str(tmp_path), samples, per_file=4)            # 3 files: 4 + 4 + 2 samples
    md = ModelDescription(g["model_json"], sample_dimensions(samples[0]))
    eng = _SpecEngine(md)
    out_name = md.get_output_info()[0]
    feats = [f[0] for f in eng.features]
    scale = {"traffic": lambda v: v * 0.5}
    batches = list(NativeIngest.batches_parallel(eng, str(tmp_path), workers=3, label_name=out_name, feature_fns=scale))
    assert [b.n_samples for b in batches] == [4, 4, 2]
    for k, b in enumerate(batches):
        pairs = [sample_to_tensors(s, feats, out_name, md.get_adjecency_info(), [], [], True) for s in samples[4 * k:4 * k + 4]]
        want = assemble([p[0] for p in pairs], eng.entities, eng.features, eng.adjacencies, (), [p[1] for p in pairs])
        for key in want.arrays:
            ref = want.arrays[key] * np.float32(0.5) if key == "feat_traffic" else want.arrays[key]
            assert np.array_equal(b.arrays[key], ref), key


@pytest.mark.parametrize("interleave", [True, False])
def test_native_ingest_position_tables_of_multi_source_sequences(interleave):
    """Q-size (links + nodes -> paths, interleave) and the same model with an `ordered` two-source
    aggregation: the per-sample position tables of batching.position_table, bit for bit."""
    import copy
    import json
    from ignnition_b200 import synthetic
    from ignnition_b200.batching import assemble
    from ignnition_b200.generator import sample_dimensions, sample_to_tensors
    from ignnition_b200.ingest import NativeIngest
    from ignnition_b200.model_description import ModelDescription
    g = load_golden("qsize_nsfnet")
    mj = copy.deepcopy(g["model_json"])
    if not interleave:
        mj["message_passing"]["stages"][0]["stage_mp"][0]["aggregation"] = {"type": "ordered"}
    samples = [synthetic.routenet_sample("nsfnet", s, s, qsize=True) for s in range(4)] + list((g.get("samples") or [])[:1])
    md = ModelDescription(mj, sample_dimensions(samples[0]))
    eng = _SpecEngine(md)
    assert len(eng.sequences) == 1 and eng.sequences[0].interleave == interleave
    out_name = md.get_output_info()[0]
    feats = [f[0] for f in eng.features]
    pairs = [sample_to_tensors(s, feats, out_name, md.get_adjecency_info(), md.get_interleave_tensors(), [], True)
             for s in samples]
    want = assemble([p[0] for p in pairs], eng.entities, eng.features, eng.adjacencies, eng.sequences, [p[1] for p in pairs])
    ing = NativeIngest(eng, label_name=out_name)
    assert ing.parse(json.dumps(samples)) == len(samples)
    got = ing.batch()
    assert set(got.arrays) == set(want.arrays)
    for k in want.arrays:
        assert got.arrays[k].dtype == want.arrays[k].dtype, k
        assert np.array_equal(got.arrays[k], want.arrays[k]), k


# ------------------------------------------------------------------ round 2: host logic of the new pieces
def test_tf_checkpoint_bundle_round_trip(tmp_path):
    """The TensorFlow tensor-bundle reader against bundles written in the same format (several data blocks,
    prefix 'model.ckpt-N'); optimizer slots are skipped, names match by suffix + shape."""
    from ignnition_b200 import tf_checkpoint as tfc
    rng = np.random.RandomState(0)
    tensors = {"comnet_model/path_update/kernel": rng.randn(32, 96).astype(np.float32),
               "comnet_model/path_update/recurrent_kernel": rng.randn(32, 96).astype(np.float32),
               "comnet_model/path_update/bias": rng.randn(2, 96).astype(np.float32),
               "comnet_model/path_update/kernel/Adam": rng.randn(32, 96).astype(np.float32),
               "comnet_model/path_update/kernel/Adam_1": rng.randn(32, 96).astype(np.float32),
               "global_step": np.asarray(1234, np.int64),
               "scalar_f64": np.asarray(2.5, np.float64)}
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.write(prefix, tensors, block_entries=3)
    assert tfc.is_tf_checkpoint(prefix) and tfc.latest(str(tmp_path)) == prefix
    got = tfc.read(prefix)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)

    class FakeEngine:
        param_table = {"path_update/kernel": (0, (32, 96)), "path_update/recurrent_kernel": (0, (32, 96)),
                       "path_update/bias": (0, (2, 96))}
    w, step = tfc.read_for_engine(prefix, FakeEngine)
    assert step == 1234 and np.array_equal(w["path_update/kernel"], tensors["comnet_model/path_update/kernel"])
    FakeEngine.param_table = dict(FakeEngine.param_table, **{"link_update/kernel": (0, (32, 96))})
    with pytest.raises(RuntimeError, match="IGNNITION: the checkpoint"):
        tfc.read_for_engine(prefix, FakeEngine)
    with pytest.raises(RuntimeError, match="bad magic"):
        bad = tmp_path / "x.index"
        bad.write_bytes(b"\0" * 64)
        tfc.read_index(str(bad))


def test_tf_checkpoint_snappy_and_prefix_compression():
    """leveldb block decoding with shared key prefixes, and the snappy decoder (literal + the three copy forms)"""
    from ignnition_b200 import tf_checkpoint as tfc
    body = b""
    for shared, key, val in ((0, b"layer/bias", b"A"), (6, b"kernel", b"BC"), (0, b"z", b"")):
        body += bytes([shared, len(key), len(val)]) + key + val
    block = body + (0).to_bytes(4, "little") + (1).to_bytes(4, "little")
    assert list(tfc._block_entries(block)) == [(b"layer/bias", b"A"), (b"layer/kernel", b"BC"), (b"z", b"")]
    # "abcdabcdabcdabcd" + "x": literal 'abcd', copy(off 4, len 12) as 2-byte-offset form, literal 'x'
    comp = bytes([17]) + bytes([3 << 2]) + b"abcd" + bytes([((12 - 1) << 2) | 2, 4, 0]) + bytes([0]) + b"x"
    assert tfc._snappy(comp) == b"abcd" * 4 + b"x"
    comp1 = bytes([8]) + bytes([3 << 2]) + b"abcd" + bytes([((4 - 4) << 2) | 1, 4])      # 1-byte-offset form
    assert tfc._snappy(comp1) == b"abcdabcd"


def test_read_dataset_shuffle_is_reproducible_across_ranks(tmp_path):
    """every rank of a data-parallel run must read the same sample stream."""
    import io
    import json
    import tarfile
    from ignnition_b200.generator import read_dataset
    for k in range(6):
        data = json.dumps([{"id": 10 * k + j} for j in range(3)]).encode()
        with tarfile.open(tmp_path / ("s%d.tar.gz" % k), "w:gz") as tar:
            info = tarfile.TarInfo("data.json")
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    a = [s["id"] for s in read_dataset(str(tmp_path), True, seed=5)]
    b = [s["id"] for s in read_dataset(str(tmp_path), True, seed=5)]
    c = [s["id"] for s in read_dataset(str(tmp_path), True, seed=6)]
    plain = [s["id"] for s in read_dataset(str(tmp_path), False)]
    assert a == b and sorted(a) == sorted(plain) == plain and a != c


def test_partition_bounds_and_split_counts():
    from ignnition_b200.parallel import node_bounds, split_counts
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            b = node_bounds(n, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
            sizes = [b[r + 1] - b[r] for r in range(world)]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    assert split_counts([0, 3, 3, 10]) == [3, 0, 7]


def test_learning_rate_schedules_by_name():
    """tf.keras.optimizers.schedules by name (generate_model.py:802-809), TF-2.1 formulas"""
    from ignnition_b200.train import LearningRate
    lr = LearningRate({"learning_rate": 0.01})
    assert lr(0) == lr(1000) == 0.01
    e = LearningRate({"schedule": {"type": "ExponentialDecay", "initial_learning_rate": 0.001, "decay_steps": 100,
                                   "decay_rate": 0.5, "staircase": "True"}})
    assert e(99) == 0.001 and abs(e(250) - 0.00025) < 1e-12
    p = LearningRate({"schedule": {"type": "PolynomialDecay", "initial_learning_rate": 0.1, "decay_steps": 100,
                                   "end_learning_rate": 0.01, "power": 2.0}})
    assert abs(p(0) - 0.1) < 1e-12 and abs(p(50) - (0.09 * 0.25 + 0.01)) < 1e-12 and abs(p(500) - 0.01) < 1e-12
    c = LearningRate({"schedule": {"type": "PiecewiseConstantDecay", "boundaries": [10, 20], "values": [1.0, 0.5, 0.1]}})
    assert (c(10), c(11), c(20), c(21)) == (1.0, 0.5, 0.5, 0.1)
    i = LearningRate({"schedule": {"type": "InverseTimeDecay", "initial_learning_rate": 0.1, "decay_steps": 10,
                                   "decay_rate": 1.0}})
    assert abs(i(10) - 0.05) < 1e-12
    with pytest.raises(RuntimeError, match="not built"):
        LearningRate({"schedule": {"type": "CosineDecay"}})


@pytest.mark.parametrize("case,qsize", [("routenet_nsfnet", False), ("qsize_nsfnet", True)])
def test_batch_window_equals_assembling_the_window(case, qsize):
    import json
    """Batch.take / take_rows (a dataset file parsed once by the native ingest, cut into training batches) gives the
    arrays assemble() gives for the same samples: entity offsets, shifted edge indices, position tables, labels"""
    from ignnition_b200 import synthetic
    from ignnition_b200.batching import assemble
    from ignnition_b200.generator import sample_dimensions, sample_to_tensors
    from ignnition_b200.ingest import NativeIngest
    g = load_golden(case)
    samples = [synthetic.routenet_sample("nsfnet", s, s, qsize=qsize) for s in range(9)]
    md = ModelDescription(g["model_json"], sample_dimensions(samples[0]))
    eng = _SpecEngine(md)
    out_name = md.get_output_info()[0]
    out_entity = [o for o in md.get_readout_operations() if o.type == "predict"][0].input[0]
    ing = NativeIngest(eng, label_name=out_name)
    assert ing.parse(json.dumps(samples)) == len(samples)
    whole = ing.batch()
    feats = [f[0] for f in eng.features]
    pairs = [sample_to_tensors(s, feats, out_name, md.get_adjecency_info(), md.get_interleave_tensors(), [], True)
             for s in samples]
    for lo, hi in ((0, 9), (0, 3), (3, 4), (4, 9)):
        win = whole.take_rows(whole.take(lo, hi, eng.adjacencies), lo, hi, eng.features, out_entity)
        want = assemble([p[0] for p in pairs[lo:hi]], eng.entities, eng.features, eng.adjacencies, eng.sequences,
                        [p[1] for p in pairs[lo:hi]])
        assert win.n_samples == want.n_samples and win.num == want.num and win.n_edges == want.n_edges
        assert win.max_seq == want.max_seq and set(win.arrays) == set(want.arrays)
        for k in want.arrays:
            assert np.array_equal(win.arrays[k], want.arrays[k]), (k, lo, hi)


# ------------------------------------------------------------------ the two oracles on the keywords added in round 2
def _sum_mpnn(hidden=32, agg="sum"):
    return {
        "entities": [{"name": "node", "hidden_state_dimension": hidden, "features": [{"name": "x", "normalization": "None"}]}],
        "message_passing": {"num_iterations": 2, "stages": [{"stage_name": "s", "stage_mp": [{
            "destination_entity": "node",
            "source_entities": [{"name": "node", "adj_vector": "adj", "message": [{"type": "direct_assignation"}]}],
            "aggregation": {"type": agg}, "update": {"type": "recurrent_neural_network", "nn_name": "rec"}}]}]},
        "readout": [{"type": "predict", "input": ["node"], "label": "y", "nn_name": "ro"}],
        "neural_networks": [{"nn_name": "rec", "nn_type": "recurrent_neural_network", "recurrent_type": "GRU"},
                            {"nn_name": "ro", "nn_type": "feed_forward", "nn_architecture": [
                                {"type_layer": "Dense", "units": 8, "activation": "relu"},
                                {"type_layer": "Dense", "units": 1, "activation": "None"}]}],
        "learning_options": {"loss": "MeanSquaredError", "optimizer": {"type": "Adam"}},
    }


def _sum_sample(rng, n):
    adj = {"v%d" % d: ["v%d" % s for s in rng.randint(0, n, rng.randint(1, 5))] for d in range(n)}
    return {"entities": {"v%d" % i: "node" for i in range(n)}, "adj": adj, "x": rng.randn(n, 3).tolist(),
            "y": rng.randn(n).tolist()}


def _both_oracles(mj, samples):
    from oracle.torch_port import TorchOracle
    dims = sample_dimensions(samples[0])
    md = ModelDescription(mj, dims)
    feats = [f.name for f in md.get_all_features()]
    out, _, _ = md.get_output_info()
    o64 = orc.Oracle(mj, dims, dtype=np.float64)
    w = o64.init_weights(7)
    to = TorchOracle(mj, dims)
    for s in samples:
        tens, _ = sample_to_tensors(s, feats, out, md.get_adjecency_info(), md.get_interleave_tensors(),
                                    md.get_additional_input_names(), True)
        p_n = o64.forward(tens, w)
        p_t = to.forward(tens, to.params(w)).detach().numpy()
        assert p_n.reshape(-1).shape == p_t.reshape(-1).shape
        assert np.abs(p_t.reshape(-1) - p_n.reshape(-1)).max() < 1e-12
    return md, o64, w


@pytest.mark.parametrize("agg", ["sum", "ordered"])
def test_oracles_agree_on_gru_reset_after_false(agg):
    """GRUCell(reset_after=False) from the JSON's cell parameters: one bias vector [3 units] in both oracles, the reset
    gate before the candidate's recurrent product; and the v1 step differs from the v2 step on the same numbers"""
    rng = np.random.RandomState(2)
    mj = _sum_mpnn(16, agg)
    mj["neural_networks"][0]["reset_after"] = False
    md, o64, w = _both_oracles(mj, [_sum_sample(rng, 12), _sum_sample(rng, 30)])
    assert w["node_update/bias"].shape == (48,)
    x, h = rng.randn(5, 16), rng.randn(5, 16)
    K, R = w["node_update/kernel"], w["node_update/recurrent_kernel"]
    v1 = orc.gru_cell(x, h, K, R, w["node_update/bias"], False)
    v2 = orc.gru_cell(x, h, K, R, np.stack([w["node_update/bias"], np.zeros(48)]), True)
    assert np.abs(v1 - v2).max() > 1e-3
    u = 16                                          # Keras v1 formula, written out
    mx = x @ K + w["node_update/bias"]
    z = 1 / (1 + np.exp(-(mx[:, :u] + h @ R[:, :u])))
    r = 1 / (1 + np.exp(-(mx[:, u:2 * u] + h @ R[:, u:2 * u])))
    hh = np.tanh(mx[:, 2 * u:] + (r * h) @ R[:, 2 * u:])
    assert np.abs(v1 - (z * h + (1 - z) * hh)).max() < 1e-12


def test_oracles_agree_on_dot_product_and_dropout():
    """readout `dot_product` = tf.tensordot(a, b, axes=0) registered 1 wide (generate_model.py:375-376), and Dropout
    layers (identity outside training; the default layer names keep counting them)"""
    rng = np.random.RandomState(3)
    mj = _sum_mpnn(16)
    mj["neural_networks"].append({"nn_name": "to1", "nn_type": "feed_forward", "nn_architecture": [
        {"type_layer": "Dense", "units": 1, "activation": "tanh"}]})
    mj["neural_networks"][1]["nn_architecture"].insert(1, {"type_layer": "Dropout", "rate": 0.3})
    mj["readout"] = [{"type": "neural_network", "input": ["node"], "nn_name": "to1", "output_name": "n1"},
                     {"type": "product", "type_product": "dot_product", "input": ["node", "n1"], "output_name": "cells"},
                     {"type": "predict", "input": ["cells"], "label": "y", "nn_name": "ro"}]
    samples = [_sum_sample(rng, 4), _sum_sample(rng, 7)]
    md, o64, w = _both_oracles(mj, samples)
    assert any(k.endswith("layer_2_Dense_readout/kernel") for k in w) and not any("Dropout" in k for k in w)
    ff = md.get_readout_operations()[-1].architecture
    assert ff.dropout and [l.type_layer for l in ff.layers] == ["Dense", "Dense"]


def test_oracles_agree_on_attention_over_two_sources():
    """generate_model.py:523-543: one edge list for both sources, the second source's padded columns start at ITS OWN
    edge count per destination, colliding cells add up (quirk 7)"""
    rng = np.random.RandomState(5)
    da = [{"type": "direct_assignation"}]
    upd = {"type": "recurrent_neural_network", "nn_name": "rec"}
    ents = [{"name": n, "hidden_state_dimension": 8, "features": [{"name": f, "normalization": "None"}]}
            for n, f in (("link", "cap"), ("node", "deg"), ("path", "tr"))]
    mj = {"entities": ents,
          "message_passing": {"num_iterations": 2, "stages": [
              {"stage_name": "s1", "stage_mp": [{"destination_entity": "path",
                                                 "source_entities": [{"name": "link", "adj_vector": "lp", "message": da},
                                                                     {"name": "node", "adj_vector": "np", "message": da}],
                                                 "aggregation": {"type": "attention"}, "update": upd}]},
              {"stage_name": "s2", "stage_mp": [{"destination_entity": "link",
                                                 "source_entities": [{"name": "path", "adj_vector": "pl", "message": da}],
                                                 "aggregation": {"type": "sum"}, "update": upd}]}]},
          "readout": [{"type": "predict", "input": ["path"], "label": "y", "nn_name": "ro"}],
          "neural_networks": _sum_mpnn()["neural_networks"],
          "learning_options": {"loss": "MeanSquaredError", "optimizer": {"type": "Adam"}}}
    samples = []
    for n_link, n_node, n_path in ((5, 4, 7), (9, 6, 20)):
        lp, npth, pl = {}, {}, {}
        for p_ in range(n_path):
            ls = rng.choice(n_link, rng.randint(2, min(5, n_link) + 1), replace=False)      # more links than nodes: collisions
            lp["p%d" % p_] = ["l%d" % l for l in ls]
            npth["p%d" % p_] = ["n%d" % n for n in rng.choice(n_node, rng.randint(1, 3), replace=False)]
            for l in ls:
                pl.setdefault("l%d" % l, []).append("p%d" % p_)
        ent = {**{"l%d" % i: "link" for i in range(n_link)}, **{"n%d" % i: "node" for i in range(n_node)},
               **{"p%d" % i: "path" for i in range(n_path)}}
        samples.append({"entities": ent, "lp": lp, "np": npth, "pl": pl, "cap": rng.rand(n_link).tolist(),
                        "deg": rng.rand(n_node).tolist(), "tr": rng.rand(n_path).tolist(), "y": rng.rand(n_path).tolist()})
    _both_oracles(mj, samples)


# ------------------------------------------------------------------ kernel choices of the Engine, without a device
def test_engine_plan_and_one_launch_gating_on_the_host():
    """Engine(model, device="plan") compiles the description into kernel choices and the parameter table without a GPU.
    The one-launch loop of small graphs (csrc/small_graph.cu) is offered for models made of ordered / sum updates of
    16 / 32-wide source states, and only for row lengths it was verified on: walks of at most one block of entries
    (units per lane group), sums of any length at 32 units (longer rows take the per-stage kernels)."""
    from ignnition_b200 import Engine
    for case, ok in (("routenet_nsfnet", True), ("qsize_nsfnet", True), ("routenet_geant2", True)):
        g = load_golden(case)
        eng = Engine(ModelDescription(g["model_json"], g["reference_meta"]["dimensions"]), device="plan")
        assert eng.device is None and not hasattr(eng, "weights") and eng._n_params > 0
        assert eng._small_program_ok() == ok
        batch = eng.assemble([orc.normalize_inputs(g["model_json"], t) for t in g["reference_tensors"][:2]])
        assert eng._small_fit(batch.max_seq)                          # paths of <= 5 links, any number of paths per link
        walk = [p for st in eng.plans for p in st if p.kind == "seq_gru"][0]
        longer = dict(batch.max_seq)
        longer[walk.adjs[0].name] = 33
        assert not eng._small_fit(longer)
        assert not eng._small_fit({})                                 # unknown lengths: never
    # 16-wide states: sums are limited to one block as well; 64-wide or attention models never take the loop
    mj = _sum_mpnn(16)
    eng = Engine(ModelDescription(mj, {"x": 3, "adj": 0}), device="plan")
    assert eng._small_program_ok() and eng._small_fit({"adj": 16}) and not eng._small_fit({"adj": 17})
    eng32 = Engine(ModelDescription(_sum_mpnn(32), {"x": 3, "adj": 0}), device="plan")
    assert eng32._small_fit({"adj": 500})
    for bad in (_sum_mpnn(64), _sum_mpnn(32, "attention"), _sum_mpnn(32, "mean")):
        assert not Engine(ModelDescription(bad, {"x": 3, "adj": 0}), device="plan")._small_program_ok()
    v1 = _sum_mpnn(32)
    v1["neural_networks"][0]["reset_after"] = False
    e1 = Engine(ModelDescription(v1, {"x": 3, "adj": 0}), device="plan")
    assert not e1._small_program_ok() and e1.param_table["node_update/bias"][1] == (96,)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        Engine(ModelDescription(_sum_mpnn(32), {"x": 3, "adj": 0}), device="cpu")
