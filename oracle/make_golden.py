"""Generate ``tests/golden/*.json`` (ORACLE / test infrastructure).

Run in the build container only (needs ``/root/reference``):

    python oracle/make_golden.py

The INTEGER fixtures come from the reference itself: ``code/utils/generator_std_to_framework.py``
and ``code/utils/json_operations.py`` are imported from /root/reference with a stub ``tensorflow``
module (the generator uses TF only for logging, the parser only for ``regularizers.l2``) and run on
seeded synthetic samples written in the reference's on-disk format.  The FLOAT fixtures are produced
by the fp64 run of ``oracle/ignnition_oracle.py`` (TensorFlow 2.1 cannot be installed here: parity
of the float half is unpinned against TF, see the oracle's header).
"""

import json
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from ignnition_b200 import synthetic  # noqa: E402  (only the dataset writer / sample generator)
from oracle.ignnition_oracle import Oracle, normalize_inputs  # noqa: E402


def import_reference():
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.keras.activations", "tensorflow.keras.losses",
                 "tensorflow.keras.optimizers", "tensorflow.keras.optimizers.schedules", "keras",
                 "keras.backend"):
        sys.modules[name] = MagicMock()
    sys.path.insert(0, os.path.join(REF, "code", "utils"))
    import generator_std_to_framework as gen
    import json_operations as jo
    return gen, jo


HAND = {   # 3 paths / 3 links / 3 nodes, Q-size layout, written by hand
    "traffic": [0.2, 0.3, 0.4], "delay": [0.1, 0.2, 0.3], "jitter": [1.0, 1.0, 1.0],
    "link_capacity": [10.0, 25.0, 40.0], "queue_sizes": [8.0, 16.0, 32.0],
    "entities": {"l0": "link", "n0": "node", "p0": "path", "l1": "link", "n1": "node", "p1": "path",
                 "l2": "link", "n2": "node", "p2": "path"},
    "adj_links_paths": {"p1": ["l1", "l2"], "p0": ["l0", "l1", "l2"], "p2": ["l2"]},
    "adj_nodes_paths": {"p1": ["n1", "n2"], "p0": ["n0", "n1", "n2"], "p2": ["n2"]},
    "adj_paths_links": {"l1": ["p1", "p0"], "l2": ["p1", "p0", "p2"], "l0": ["p0"]},
    "adj_paths_nodes": {"n1": ["p1", "p0"], "n2": ["p1", "p0", "p2"], "n0": ["p0"]},
    "path_interleave": ["node", "link"],
}


def to_jsonable(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            v = v.tolist()
        elif isinstance(v, (np.integer,)):
            v = int(v)
        out[k] = v
    return out


def run_reference_generator(gen, jo, model_json_path, samples):
    """Model_information + generator of the reference on a temp dataset; returns per-sample dicts."""
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "code"))            # json_operations.py:139 opens ./utils/schema.json
    try:
        with tempfile.TemporaryDirectory() as tmp:
            synthetic.write_dataset(tmp, samples, per_file=len(samples))
            from ignnition_b200.generator import sample_dimensions
            dims = sample_dimensions(samples[0])
            info = jo.Model_information(model_json_path, dims)
            feats = [f.name.encode() for f in info.get_all_features()]
            adj = [[x.encode() for x in a] for a in info.get_adjecency_info()]
            inter = [[x.encode() for x in a] for a in info.get_interleave_tensors()]
            out_name, out_norm, out_denorm = info.get_output_info()
            add = [a.encode() for a in info.get_additional_input_names() if a.encode() not in feats]
            got = list(gen.generator(tmp.encode(), feats, out_name.encode(), adj, inter, add, True))
            meta = {
                "adjacency_info": info.get_adjecency_info(),
                "interleave_tensors": info.get_interleave_tensors(),
                "interleave_sources": info.get_interleave_sources(),
                "features": [[f.name, f.size, f.normalization] for f in info.get_all_features()],
                "output_info": [out_name, out_norm, out_denorm],
                "mp_iterations": info.get_mp_iterations(),
                "input_dimensions": info.get_input_dimensions(),
                "additional_input": info.get_additional_input_names(),
                "loss": info.get_loss(),
                "optimizer": info.get_optimizer(),
                "stages": [[name, [[mp.destination_entity, [s.name for s in mp.source_entities],
                                    mp.aggregation.type, mp.update.type] for mp in mps]]
                           for name, mps in info.get_mp_instances()],
                "dimensions": dims,
            }
    finally:
        os.chdir(cwd)
    return got, meta


def float_golden(model_json, dims, tensors, seed=1234):
    orc = Oracle(model_json, dims, dtype=np.float64)
    w = orc.init_weights(seed)
    w32 = {k: v.astype(np.float32) for k, v in w.items()}           # weights are fp32 values
    tensors = normalize_inputs(model_json, tensors)                  # float32 features, as input_fn
    pred, st = orc.forward(tensors, w32, return_states=True)
    return {"weight_seed": seed, "predictions_fp64": pred.reshape(-1).tolist(),
            "state_checksums_fp64": {k: float(np.asarray(v, dtype=np.float64).sum()) for k, v in st.items()},
            "state_abs_checksums_fp64": {k: float(np.abs(np.asarray(v, dtype=np.float64)).sum())
                                          for k, v in st.items()}}


def main():
    gen, jo = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rn_json = os.path.join(REF, "examples", "Routenet", "model_description.json")
    qs_json = os.path.join(REF, "examples", "Q-size", "model_description.json")

    cases = {
        "routenet_nsfnet": (rn_json, [synthetic.routenet_sample("nsfnet", 0, 0),
                                      synthetic.routenet_sample("nsfnet", 1, 1)]),
        "qsize_hand": (qs_json, [HAND]),
        "qsize_nsfnet": (qs_json, [synthetic.routenet_sample("nsfnet", 0, 0, qsize=True)]),
        "routenet_geant2": (rn_json, [synthetic.routenet_sample("geant2", 0, 0)]),
        # BASELINE config 4 shape (50 nodes, ~2.4 k paths): the training benchmark's topology
        "routenet_synth50": (rn_json, [synthetic.routenet_sample("synth50", 0, 0)]),
    }
    recipes = {"routenet_geant2": ["geant2", 0, 0], "routenet_synth50": ["synth50", 0, 0]}   # samples too big to store
    for name, (mj, samples) in cases.items():
        got, meta = run_reference_generator(gen, jo, mj, samples)
        assert len(got) == len(samples), (name, len(got))
        model_json = json.load(open(mj))
        fixture = {
            "generated_by": "oracle/make_golden.py: reference generator_std_to_framework.generator + "
                            "json_operations.Model_information under a stub tensorflow",
            "model": os.path.relpath(mj, REF),
            "model_json": model_json,
            "reference_meta": meta,
            "samples": samples if name not in recipes else None,
            "sample_recipe": recipes.get(name),
            "reference_tensors": [to_jsonable(d) for d, _ in got],
            "reference_labels": [list(map(float, y)) for _, y in got],
        }
        if name not in recipes:
            fixture["oracle_float"] = [float_golden(model_json, meta["dimensions"], d) for d, _ in got]
        path = os.path.join(out_dir, name + ".json")
        with open(path, "w") as fh:
            json.dump(fixture, fh, separators=(",", ":"))
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
