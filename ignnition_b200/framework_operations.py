"""The reference's public entry points on the B200 engine: ``create_model``, ``train_and_evaluate``,
``predict``, ``debug`` (reference ``code/utils/framework_operations.py:42-268``, ``readme.md:77-95``).

Same files drive it: ``./train_options.ini`` ([PATHS] train_dataset, eval_dataset, predict_dataset,
warm_start_path, json_path, model_dir; [TRAINING_OPTIONS] batch_size, train_steps, shuffle_*,
eval_samples, save_checkpoints_secs, keep_checkpoint_max, throttle_secs, execute_gpu -- reference
``code/train_options.ini:1-50``), the same ``model_description.json`` and the user's ``main.py`` with
its normalisation functions (looked up by name, ``generate_model.py:68,77``).  This is host glue:
the tf.estimator / tf.data plumbing of the reference is NOT rebuilt; checkpoints are ``.npz`` files
of the named variables (Keras layouts), written every ``save_checkpoints_secs``.
"""

from __future__ import annotations

import configparser
import glob
import itertools
import os
import sys
import tarfile
import time
from typing import Callable, Dict, Iterator, List, Optional

import numpy as np

from .generator import find_dataset_dimensions, read_dataset, sample_to_tensors
from .model_description import ModelDescription

CONFIG: Optional[configparser.ConfigParser] = None


def _log(msg: str):
    print("IGNNITION: " + msg, file=sys.stderr, flush=True)


def load_config(path: str = "./train_options.ini") -> configparser.ConfigParser:
    """ExtendedInterpolation ini, read relative to the CWD like the reference (:34-36)."""
    global CONFIG
    cfg = configparser.ConfigParser()
    cfg._interpolation = configparser.ExtendedInterpolation()
    if not cfg.read(path):
        raise RuntimeError("IGNNITION: the configuration file %s was not found" % path)
    CONFIG = cfg
    return cfg


def _cfg() -> configparser.ConfigParser:
    return CONFIG if CONFIG is not None else load_config()


def create_model(config_path: Optional[str] = None) -> ModelDescription:
    cfg = load_config(config_path) if config_path else _cfg()
    dims = find_dataset_dimensions(cfg["PATHS"]["train_dataset"])
    return ModelDescription(cfg["PATHS"]["json_path"], dims)


# ------------------------------------------------------------------ normalisation (generate_model.py:46-86)
def _resolve(name: Optional[str], namespace: Optional[dict]) -> Optional[Callable]:
    if name is None or str(name) == "None":
        return None
    for ns in (namespace, getattr(sys.modules.get("__main__"), "__dict__", None),
               getattr(sys.modules.get("main"), "__dict__", None)):
        if ns and name in ns and callable(ns[name]):
            return ns[name]
    return None


def normalize(model: ModelDescription, x: dict, y=None, namespace: Optional[dict] = None):
    for f in model.get_all_features():
        fn = _resolve(f.normalization, namespace)
        v = np.asarray(x[f.name], dtype=np.float32)
        if str(f.normalization) != "None":
            if fn is None:
                raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file."
                                   % f.normalization)
            v = np.asarray(fn(v, f.name), dtype=np.float32)
        x[f.name] = v
    if y is None:
        return x
    out_name, out_norm, _ = model.get_output_info()
    y = np.asarray(y, dtype=np.float32)
    if str(out_norm) != "None":
        fn = _resolve(out_norm, namespace)
        if fn is None:
            raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file." % out_norm)
        y = np.asarray(fn(y, out_name), dtype=np.float32)
    return x, y


def samples_of(model: ModelDescription, directory: str, training: bool, shuffle: bool = False,
               namespace: Optional[dict] = None, repeat: bool = False, seed: Optional[int] = None) -> Iterator:
    """The reference's input_fn as a Python iterator of normalised tensor dicts (+ labels).  ``seed``: the file
    shuffle of pass k uses seed + k, so that all ranks of a data-parallel run read the same stream."""
    feats = [f.name for f in model.get_all_features()]
    out_name, _, _ = model.get_output_info()
    adj, inter = model.get_adjecency_info(), model.get_interleave_tensors()
    extra = [a for a in model.get_additional_input_names() if a not in feats]
    epoch = 0
    while True:
        n = 0
        for sample in read_dataset(directory, shuffle, None if seed is None else seed + epoch):
            n += 1
            if training:
                x, y = sample_to_tensors(sample, feats, out_name, adj, inter, extra, True)
                yield normalize(model, x, y, namespace)
            else:
                yield normalize(model, sample_to_tensors(sample, feats, out_name, adj, inter, extra, False),
                                namespace=namespace)
        if not repeat or n == 0:
            return
        epoch += 1


def native_batches(model: ModelDescription, engine, directory: str, training: bool,
                   namespace: Optional[dict] = None, workers: int = 8) -> Iterator:
    """One normalised :class:`Batch` per ``*.tar.gz`` of ``directory`` through the C++ ingest
    (``ingest.NativeIngest``): the same arrays as ``samples_of`` + ``Engine.assemble`` without the per-edge
    Python loops.  The user's normalisation functions are applied to the whole batch array at once, so
    they must be elementwise (every one in ``examples/`` is)."""
    from .ingest import NativeIngest
    fns = {}
    for f in model.get_all_features():
        if str(f.normalization) != "None":
            fn = _resolve(f.normalization, namespace)
            if fn is None:
                raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file."
                                   % f.normalization)
            fns[f.name] = (lambda v, fn=fn, name=f.name: fn(v, name))
    out_name, out_norm, _ = model.get_output_info()
    label_fn = None
    if training and str(out_norm) != "None":
        fn = _resolve(out_norm, namespace)
        if fn is None:
            raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file." % out_norm)
        label_fn = lambda v, fn=fn: fn(v, out_name)
    return NativeIngest.batches_parallel(engine, directory, workers, out_name if training else None,
                                         feature_fns=fns, label_fn=label_fn)


def native_train_batches(model: ModelDescription, engine, directory: str, batch_size: int,
                         namespace: Optional[dict] = None, rank: int = 0, world: int = 1, shuffle: bool = False,
                         seed: int = 0) -> Iterator:
    """Endless stream of labelled training batches of ``batch_size`` samples through the C++ ingest: every
    ``*.tar.gz`` is parsed ONCE into one big batch (a background thread parses the next file while this one is
    consumed) and cut into windows with ``Batch.take``.  Data-parallel runs shard the FILES: rank r reads files
    r, r + world, ... (all files when there are fewer files than ranks, then the windows are dealt round-robin), so no
    rank parses text another rank uses.  Replaces the per-sample Python generator of the reference's input_fn
    (generate_model.py:102-198) in the training loop."""
    import random
    from concurrent.futures import ThreadPoolExecutor

    from .ingest import NativeIngest
    paths = sorted(glob.glob(os.path.join(directory, "*.tar.gz")))
    if not paths:
        raise RuntimeError("IGNNITION: no *.tar.gz dataset files in %s" % directory)
    deal = len(paths) < world                       # too few files: every rank reads all, windows are dealt out
    mine = paths if deal else paths[rank::world]
    out_name, out_norm, _ = model.get_output_info()
    out_entity = [o for o in model.get_readout_operations() if o.type == "predict"][0].input[0]
    fns = {}
    for f in model.get_all_features():
        if str(f.normalization) != "None":
            fn = _resolve(f.normalization, namespace)
            if fn is None:
                raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file."
                                   % f.normalization)
            fns[f.name] = (lambda v, fn=fn, name=f.name: fn(v, name))
    label_fn = None
    if str(out_norm) != "None":
        lf = _resolve(out_norm, namespace)
        if lf is None:
            raise RuntimeError("IGNNITION: The normalization function %s is not defined in the main file." % out_norm)
        label_fn = lambda v: lf(v, out_name)

    def parse(path):
        ing = NativeIngest(engine, out_name)
        with tarfile.open(path, "r:gz") as tar:
            ing.parse(tar.extractfile("data.json").read())
        return ing.batch(feature_fns=fns, label_fn=label_fn)

    epoch, k = 0, 0
    with ThreadPoolExecutor(max_workers=1) as pool:
        order = list(mine)
        if shuffle:
            random.Random(seed).shuffle(order)
        nxt = pool.submit(parse, order[0])
        i = 0
        while True:
            whole = nxt.result()
            i += 1
            if i == len(order):
                i, epoch = 0, epoch + 1
                order = list(mine)
                if shuffle:
                    random.Random(seed + epoch).shuffle(order)
            nxt = pool.submit(parse, order[i])
            for lo in range(0, whole.n_samples, batch_size):
                k += 1
                if deal and (k - 1) % world != rank:
                    continue
                hi = min(lo + batch_size, whole.n_samples)
                yield whole.take_rows(whole.take(lo, hi, engine.adjacencies), lo, hi, engine.features, out_entity)


def eval_metrics(labels: np.ndarray, preds: np.ndarray) -> Dict[str, float]:
    """label/prediction mean, MAE, MRE, R^2 (generate_model.py:770-787, 201-216)."""
    y, p = labels.astype(np.float64).reshape(-1), preds.astype(np.float64).reshape(-1)
    return {"label/mean": float(y.mean()), "prediction/mean": float(p.mean()),
            "mae": float(np.abs(y - p).mean()), "mre": float((np.abs(y - p) / np.abs(y)).mean()),
            "r-squared": float(1.0 - ((y - p) ** 2).sum() / ((y - y.mean()) ** 2).sum())}


def save_checkpoint(engine, trainer, model_dir: str, keep_max: int) -> str:
    os.makedirs(model_dir, exist_ok=True)
    path = os.path.join(model_dir, "model.ckpt-%d.npz" % trainer.step)
    np.savez(path, __step__=np.int64(trainer.step), __adam_m__=trainer.m.cpu().numpy(),
             __adam_v__=trainer.v.cpu().numpy(), **engine.get_weights())
    old = sorted(glob.glob(os.path.join(model_dir, "model.ckpt-*.npz")), key=_ckpt_step)
    for p in old[:-keep_max] if keep_max > 0 else []:
        os.remove(p)
    return path


def _ckpt_step(path: str) -> int:
    name = os.path.basename(path)
    try:
        return int(name[len("model.ckpt-"):].split(".")[0])
    except ValueError:
        return -1


def load_checkpoint(engine, path: str, trainer=None) -> int:
    """Warm start: every stored kernel / recurrent_kernel / bias variable (framework_operations.py:126-129).
    With ``trainer`` the Adam moments and the step are restored too (a resume, not only a warm start).
    Reads this engine's ``.npz`` checkpoints and TensorFlow checkpoints of the reference (``model.ckpt-N.index``
    + ``.data-*``: ``tf_checkpoint.read``), whose variables map 1:1 by name and layout."""
    from . import tf_checkpoint
    if os.path.isdir(path):
        files = sorted(glob.glob(os.path.join(path, "model.ckpt-*.npz")), key=_ckpt_step)
        if files:
            path = files[-1]
        else:
            prefix = tf_checkpoint.latest(path)
            if prefix is None:
                raise RuntimeError("IGNNITION: no checkpoint found in " + path)
            path = prefix
    if tf_checkpoint.is_tf_checkpoint(path):
        weights, step = tf_checkpoint.read_for_engine(path, engine)
        engine.set_weights(weights)
        if trainer is not None:
            trainer.step = step
        return step
    data = np.load(path)
    engine.set_weights({k: data[k] for k in data.files if not k.startswith("__")})
    step = int(data["__step__"]) if "__step__" in data.files else 0
    if trainer is not None:
        trainer.step = step
        if "__adam_m__" in data.files:
            import torch
            trainer.m.copy_(torch.from_numpy(data["__adam_m__"]))
            trainer.v.copy_(torch.from_numpy(data["__adam_v__"]))
    return step


def evaluate(model: ModelDescription, engine, directory: str, n_samples: int, shuffle: bool = False,
             namespace: Optional[dict] = None, native_ingest: Optional[bool] = None) -> Dict[str, float]:
    """``native_ingest`` (default: the IGNNITION_NATIVE_INGEST environment variable, off): read the dataset
    through the C++ ingest, one batch per file, instead of the per-sample Python generator."""
    out_name, _, out_denorm = model.get_output_info()
    ys, ps = [], []
    if native_ingest is None:
        native_ingest = os.environ.get("IGNNITION_NATIVE_INGEST", "0") not in ("", "0")
    if native_ingest:
        from .ingest import NativeIngest
        native_ingest = NativeIngest.supported(engine) and not shuffle
    if native_ingest:
        n_seen = 0
        for batch in native_batches(model, engine, directory, True, namespace):
            graph = engine.build_graph(engine.upload(batch))
            ps.append(engine.forward(graph).cpu().numpy().reshape(-1))
            ys.append(batch.arrays["labels"].reshape(-1))
            n_seen += batch.n_samples
            if n_samples and n_seen >= n_samples:
                break
    it = samples_of(model, directory, True, shuffle, namespace) if not native_ingest else iter(())
    while True:
        chunk = list(itertools.islice(it, min(64, n_samples - len(ys)) if n_samples else 64))
        if not chunk:
            break
        pred = engine.forward(engine.prepare([c[0] for c in chunk])).cpu().numpy().reshape(-1)
        ps.append(pred)
        ys += [c[1].reshape(-1) for c in chunk]
        if n_samples and len(ys) >= n_samples:
            break
    y, p = np.concatenate(ys), np.concatenate(ps)
    loss = float(np.mean((y - p) ** 2))
    fn = _resolve(out_denorm, namespace)
    if fn is not None:
        y, p = np.asarray(fn(y, out_name)), np.asarray(fn(p, out_name))
    else:
        _log("A denormalization function for output %s was not defined. The output (and statistics) will "
             "use the normalized values." % out_name)
    return dict(eval_metrics(y, p), loss=loss)


def train_and_evaluate(model: ModelDescription, namespace: Optional[dict] = None, max_steps: Optional[int] = None,
                       engine=None):
    """tf.estimator.train_and_evaluate of the reference (:108-166) on Engine + Trainer."""
    import datetime

    import torch

    from .engine import Engine
    from .parallel import rank_world
    from .train import Trainer
    cfg = _cfg()
    opt = cfg["TRAINING_OPTIONS"]
    rank, world, local = rank_world()
    _log("Starting the training and evaluation process...")
    torch.cuda.set_device(local)
    engine = engine or Engine(model, device=torch.device("cuda", local))
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=engine.device)
    trainer = Trainer(engine, world_size=world)
    if cfg.has_option("PATHS", "warm_start_path") and cfg["PATHS"]["warm_start_path"].strip():
        load_checkpoint(engine, cfg["PATHS"]["warm_start_path"],
                        trainer if opt.get("resume", "False") == "True" else None)
    model_dir = os.path.join(cfg["PATHS"]["model_dir"], "experiment_" + str(datetime.datetime.now()).replace(" ", "_"))
    batch = int(opt["batch_size"])
    steps = int(opt["train_steps"]) if max_steps is None else max_steps
    keep = int(opt.get("keep_checkpoint_max", 20))
    ckpt_secs, eval_secs = float(opt.get("save_checkpoints_secs", 300)), float(opt.get("throttle_secs", 300))
    # every rank reads the SAME stream (file shuffle seeded identically on all ranks) and keeps its slice of each
    # global batch; the global prediction count is checked across ranks below
    native = opt.get("native_ingest", os.environ.get("IGNNITION_NATIVE_INGEST", "False")) in ("True", "1", "true")
    shuffle = opt.get("shuffle_train_samples", "True") == "True"
    if native:
        # data.json text -> batch arrays in C++, one file ahead, files sharded over the ranks (native_train_batches)
        stream = native_train_batches(model, engine, cfg["PATHS"]["train_dataset"], batch, namespace, rank, world,
                                      shuffle, int(opt.get("shuffle_seed", 0)))
        it = None
    else:
        it = samples_of(model, cfg["PATHS"]["train_dataset"], True, shuffle, namespace, repeat=True,
                        seed=int(opt.get("shuffle_seed", 0)) if world > 1 else None)
    t_ckpt = t_eval = time.time()
    history = []
    for step in range(steps):
        if native:
            b = next(stream)
            n_loc = int(b.arrays["labels"].size)
            n_glob = n_loc
            if world > 1:                         # ranks read different files: the global count is their sum
                cnt = torch.tensor([n_loc], dtype=torch.int64, device=engine.device)
                torch.distributed.all_reduce(cnt)
                n_glob = int(cnt.item())
            trainer.step_batch(b, global_n=n_glob)     # repeated small shapes replay a captured graph
            chunk = mine = None
        else:
            chunk = list(itertools.islice(it, batch * world))
            if not chunk:
                break
            mine = chunk[rank::world] if world > 1 else chunk
            n_glob = sum(c[1].size for c in chunk)
        if not native and world > 1 and step % 100 == 0:         # the 1/N of the loss must be the same number on every rank
            chk = torch.tensor([n_glob, -n_glob], dtype=torch.int64, device=engine.device)
            torch.distributed.all_reduce(chk, op=torch.distributed.ReduceOp.MAX)
            if int(chk[0]) != n_glob or int(-chk[1]) != n_glob:
                raise RuntimeError("IGNNITION: the ranks of the data-parallel run read different sample streams "
                                   "(global batch of %d predictions here, %d..%d across ranks)"
                                   % (n_glob, int(-chk[1]), int(chk[0])))
        if not native:
            trainer.step_batch(engine.assemble([c[0] for c in mine], [c[1] for c in mine]), global_n=n_glob)
        if step % 10 == 0:                        # LoggingTensorHook every 10 iterations (:820-824)
            l = trainer.losses()
            history.append((step, l["loss"]))
            if rank == 0:
                _log("step %d  Loss = %.6g, Regularization loss = %.6g, Total loss = %.6g"
                     % (step, l["loss"], l["regularization_loss"], l["total_loss"]))
        now = time.time()
        if rank == 0 and (now - t_ckpt >= ckpt_secs or step == steps - 1):
            save_checkpoint(engine, trainer, model_dir, keep)
            t_ckpt = now
        if rank == 0 and (now - t_eval >= eval_secs or step == steps - 1) and cfg["PATHS"].get("eval_dataset", "").strip():
            m = evaluate(model, engine, cfg["PATHS"]["eval_dataset"], int(opt.get("eval_samples", 100)),
                         opt.get("shuffle_eval_samples", "False") == "True", namespace)
            _log("evaluation at step %d: %s" % (step, ", ".join("%s = %.6g" % kv for kv in m.items())))
            t_eval = now
    return engine, trainer, history


def predict(model: ModelDescription, namespace: Optional[dict] = None, engine=None) -> List[np.ndarray]:
    """One pass over predict_dataset with the warm-start weights (:169-236); the reference's loop never
    terminates on a finite dataset, this one stops at the end."""
    import torch

    from .engine import Engine
    cfg = _cfg()
    _log("Starting to make the predictions...")
    if not cfg.has_option("PATHS", "warm_start_path"):
        _log("The path of the model to use for the predictions is unspecified. Please add a field "
             "warm_start_path in the train_options.ini with the corresponding path to the model you want to restore.")
        sys.exit(0)
    if not cfg.has_option("PATHS", "predict_dataset"):
        _log("The path of dataset to use for the prediction is unspecified. Please add a field predict_dataset "
             "in the train_config.ini file with the corresponding path to the dataset you want to predict.")
        sys.exit(0)
    engine = engine or Engine(model, device=torch.device("cuda", torch.cuda.current_device()))
    load_checkpoint(engine, cfg["PATHS"]["warm_start_path"])
    out_name, _, out_denorm = model.get_output_info()
    fn = _resolve(out_denorm, namespace)
    if fn is None:
        _log("A denormalization function for output %s was not defined. The output will be normalized." % out_name)
    result = []
    it = samples_of(model, cfg["PATHS"]["predict_dataset"], False, False, namespace)
    out_entity = [o for o in model.get_readout_operations() if o.type == "predict"][0].input[0]
    chunk_size = int(cfg["TRAINING_OPTIONS"].get("predict_batch", 256)) if cfg.has_section("TRAINING_OPTIONS") else 256
    while True:                                   # block-diagonal batches: one forward per chunk, split per sample
        chunk = list(itertools.islice(it, chunk_size))
        if not chunk:
            break
        # chunks of one shape (the usual case: every sample of a dataset has the topology's sizes) replay one captured
        # CUDA graph: a host -> device copy and one launch per chunk instead of ~50 launches
        p = engine.forward_graphed(engine.assemble(chunk)).cpu().numpy().reshape(-1)
        off = 0
        for x in chunk:
            n = int(x["num_" + out_entity])
            q = p[off:off + n]
            off += n
            result.append(np.asarray(fn(q, out_name)) if fn is not None else q)
    return result


def debug(model: ModelDescription, engine=None) -> str:
    """The reference dumps the TF graph for TensorBoard (:239-268); here: the compiled kernel plan."""
    import torch

    from .engine import Engine
    cfg = _cfg()
    engine = engine or Engine(model, device=torch.device("cuda", torch.cuda.current_device()))
    lines = ["ignnition_b200 kernel plan", "T = %d" % engine.T]
    for stage in engine.plans:
        for p in stage:
            lines.append("  %s <- %s : %s (%s)" % (p.dst, [a.src for a in p.adjs], p.kind, p.mp.aggregation.type))
    for name, (off, shape) in engine.param_table.items():
        lines.append("  var %-60s %s @%d" % (name, shape, off))
    text = "\n".join(lines)
    out = os.path.join(cfg["PATHS"].get("debug_dir", "../"), "debug_model")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "plan.txt"), "w") as fh:
        fh.write(text + "\n")
    _log("The debug model has been generated.")
    return text
