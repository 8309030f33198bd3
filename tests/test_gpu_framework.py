"""GPU test of the reference's public flow (create_model -> debug -> train_and_evaluate -> predict) on a
synthetic dataset in the reference's on-disk format, driven by a train_options.ini."""

import os

import numpy as np
import pytest

from conftest import load_golden
from ignnition_b200 import synthetic

pytestmark = pytest.mark.gpu


def normalization_routenet(feature, feature_name):       # examples/Routenet/main.py:26-33
    if feature_name == 'traffic':
        feature = (feature - 170) / 130
    if feature_name == 'link_capacity':
        feature = (feature - 25000) / 40000
    return feature


def test_train_eval_predict_flow(tmp_path):
    import json
    from ignnition_b200 import framework_operations as ignnition, tf_shim
    tf = tf_shim
    ns = {"normalization_routenet": normalization_routenet,
          "log": lambda f, n: tf.math.log(f), "exp": lambda f, n: tf.math.exp(f)}
    g = load_golden("routenet_nsfnet")
    mj = dict(g["model_json"])
    mj["readout"] = [dict(mj["readout"][0], label_denormalization="exp")]
    (tmp_path / "model_description.json").write_text(json.dumps(mj))
    train = [synthetic.routenet_sample("nsfnet", k % 3, k) for k in range(24)]
    synthetic.write_dataset(str(tmp_path / "train"), train, per_file=8)
    synthetic.write_dataset(str(tmp_path / "eval"), train[:6], per_file=6)
    ini = tmp_path / "train_options.ini"
    ini.write_text("""[PATHS]
train_dataset: %(d)s/train
eval_dataset: %(d)s/eval
predict_dataset: %(d)s/eval
json_path: %(d)s/model_description.json
model_dir: %(d)s/CheckPoints
debug_dir: %(d)s/
[TRAINING_OPTIONS]
batch_size: 3
train_steps: 30
shuffle_train_samples: False
shuffle_eval_samples: False
eval_samples: 6
save_checkpoints_secs: 300
keep_checkpoint_max: 2
throttle_secs: 300
execute_gpu: True
""" % {"d": str(tmp_path)})
    model = ignnition.create_model(str(ini))
    assert model.get_mp_iterations() == 8
    assert "path <- ['link'] : seq_gru" in ignnition.debug(model)
    engine, trainer, history = ignnition.train_and_evaluate(model, namespace=ns)
    assert trainer.step == 30
    assert history[-1][1] < history[0][1]                       # the loss goes down
    ckpts = [f for f in os.listdir(next(p for p in (tmp_path / "CheckPoints").iterdir())) if f.endswith(".npz")]
    assert len(ckpts) == 1
    # predict restores the checkpoint and denormalises with `exp`
    ignnition.CONFIG["PATHS"]["warm_start_path"] = str(next((tmp_path / "CheckPoints").iterdir()))
    preds = ignnition.predict(model, namespace=ns)
    assert len(preds) == 6 and preds[0].shape == (182,) and np.all(preds[0] > 0)
    m = ignnition.evaluate(model, engine, str(tmp_path / "eval"), 6, namespace=ns)
    assert set(m) >= {"label/mean", "prediction/mean", "mae", "mre", "r-squared", "loss"}
    # the same evaluation with the dataset read through the C++ ingest (one batch per file)
    m2 = ignnition.evaluate(model, engine, str(tmp_path / "eval"), 6, namespace=ns, native_ingest=True)
    for k in m:
        assert abs(m2[k] - m[k]) <= 1e-5 * max(abs(m[k]), 1e-12), k


def test_training_loop_through_native_ingest(tmp_path):
    """train_and_evaluate with native_ingest: data.json text -> batch arrays in C++ (one file ahead, windows of
    batch_size samples) gives the same weights as the per-sample Python generator on the same unshuffled dataset"""
    import json
    from ignnition_b200 import framework_operations as ignnition
    ns = {"normalization_routenet": normalization_routenet, "log": lambda f, n: np.log(f), "exp": lambda f, n: np.exp(f)}
    g = load_golden("routenet_nsfnet")
    (tmp_path / "model_description.json").write_text(json.dumps(g["model_json"]))
    train = [synthetic.routenet_sample("nsfnet", k % 3, k) for k in range(24)]
    synthetic.write_dataset(str(tmp_path / "train"), train, per_file=8)         # 8 = 2 x batch_size 4: whole windows
    weights = {}
    for native in ("False", "True"):
        ini = tmp_path / ("train_options_%s.ini" % native)
        ini.write_text("""[PATHS]
train_dataset: %(d)s/train
json_path: %(d)s/model_description.json
model_dir: %(d)s/CheckPoints_%(n)s
debug_dir: %(d)s/
[TRAINING_OPTIONS]
batch_size: 4
train_steps: 9
shuffle_train_samples: False
native_ingest: %(n)s
save_checkpoints_secs: 3000
keep_checkpoint_max: 2
throttle_secs: 3000
execute_gpu: True
""" % {"d": str(tmp_path), "n": native})
        model = ignnition.create_model(str(ini))
        engine, trainer, history = ignnition.train_and_evaluate(model, namespace=ns)
        assert trainer.step == 9
        weights[native] = engine.weights.cpu().numpy().copy()
    # (weight gradients are flushed with fp32 atomics: two runs of the same data differ in the last bits)
    assert np.abs(weights["True"] - weights["False"]).max() <= 1e-5 * np.abs(weights["False"]).max()
