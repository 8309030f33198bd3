#!/usr/bin/env python
"""Benchmark of the message-passing hot path (BASELINE.json metric: RouteNet samples/sec, with the
message-passing edges/sec of every iteration reported beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one pass of the hot path over one batch of synthetic input: device adjacency build
(CSR, length order, step tables) + T message-passing iterations + readout.
 * ``value``  : whole-job samples/s with the batch's raw tensors already resident in HBM.
 * ``e2e``    : the same through the public Engine API from pinned HOST buffers: H2D of the packed
                batch, adjacency build, forward, D2H of the predictions -- all inside the timed region.
 * ``roofline``: the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak.
 * ``cpu_baseline`` / ``--impl reference``: the op-for-op CPU restatement of the reference (oracle/),
   because the reference itself needs tensorflow==2.1.0 which cannot be installed.
Weak scaling: every rank runs its own batch of the same size; no data-path collective (inference).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (golden fixture holding model json + base topology, shape, qsize, default samples per GPU)
    "routenet_geant2_b4096": ("routenet_geant2", "geant2", False, 4096),      # BASELINE config 3 (default)
    "routenet_nsfnet_b4096": ("routenet_nsfnet", "nsfnet", False, 4096),      # config 1 shape, batched
    "qsize_nsfnet_b4096": ("qsize_nsfnet", "nsfnet", True, 4096),             # config 2
    "routenet_synth50_b256": ("routenet_synth50", "synth50", False, 256),     # config 4 shape (use with --train)
}
DEFAULT_WORKLOAD = "routenet_geant2_b4096"


def load_case(name):
    fixture, shape, qsize, n = WORKLOADS[name]
    g = json.load(open(os.path.join(ROOT, "tests", "golden", fixture + ".json")))
    return g, shape, qsize, n


def feature_fns(qsize):
    # already-normalised feature ranges (examples/*/main.py normalisations applied to the raw ranges)
    if qsize:
        return {"traffic": lambda r, n: (r.uniform(0.05, 0.6, n) - 0.28) / 0.15,
                "link_capacity": lambda r, n: (r.choice([10.0, 25.0, 40.0], n) - 27.0) / 14.86,
                "queue_sizes": lambda r, n: (r.choice([1.0, 8.0, 16.0, 32.0], n) - 16.5) / 15.5}
    return {"traffic": lambda r, n: (r.uniform(40.0, 300.0, n) - 170.0) / 130.0,
            "link_capacity": lambda r, n: (r.choice([10000.0, 40000.0], n) - 25000.0) / 40000.0}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------ CPU restatement
def _cpu_worker(args):
    """Forward of `count` samples with the oracle (per-sample loop = the reference's model_fn loop)."""
    workload, count, seed = args
    import numpy as _np
    from oracle import ignnition_oracle as orc
    g, shape, qsize, _ = load_case(workload)
    dims = g["reference_meta"]["dimensions"]
    o = orc.Oracle(g["model_json"], dims, dtype=_np.float32)
    w = o.init_weights(1234)
    base = dict(g["reference_tensors"][0])
    rng = _np.random.RandomState(seed)
    fns = feature_fns(qsize)
    ent_of = {f["name"]: e["name"] for e in g["model_json"]["entities"] for f in e.get("features", [])}
    t0 = time.perf_counter()
    for _ in range(count):
        for name, fn in fns.items():
            base[name] = fn(rng, base["num_" + ent_of[name]]).astype(_np.float32)
        o.forward(base, w)
    return time.perf_counter() - t0


def cpu_samples_per_s(workload, total_samples, procs):
    import multiprocessing as mp
    per = max(1, total_samples // procs)
    t0 = time.perf_counter()
    if procs == 1:
        _cpu_worker((workload, per, 0))
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            pool.map(_cpu_worker, [(workload, per, s) for s in range(procs)])
    dt = time.perf_counter() - t0
    return per * procs / dt, per * procs, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # one worker process per core, one BLAS thread each (the matmuls are tiny: [P,32]x[32,96])
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    # calibrate on one core, then size a step to ~3 s of wall time on all cores
    _cpu_worker((args.workload, 1, 0))
    t1 = _cpu_worker((args.workload, 2, 0)) / 2
    import multiprocessing as mp
    per = max(1, min(512, int(3.0 / max(t1, 1e-4))))
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_worker, [(args.workload, max(1, per // 4), s) for s in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_worker, [(args.workload, per, 100 + k * cores + s) for s in range(cores)])
        dt = time.perf_counter() - t0
    n = per * cores
    val = n * args.steps / dt
    g, shape, qsize, _ = load_case(args.workload)
    line = {
        "impl": "reference", "metric": "routenet_samples_per_s", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "samples_per_step": n, "iterations": 8,
                   "note": "CPU restatement of the reference (oracle port); tensorflow==2.1.0 not installable"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d samples/step of %s, one process per core, per-sample forward loop "
                                   "(generate_model.py:712-724)" % (n, args.workload)},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ our arm
def measure_case(ctx, workload, n_samples, train, steps, warmup, seed_weights=0, want_e2e=True, want_kernels=True,
                 parity_samples=0, total_scale=None, graphed=False):
    """One workload through the Engine on this rank's GPU: resident-input steps, end-to-end steps from pinned host
    memory, per-kernel CUDA-event times, parity against the fp64 oracle.  Times are max over ranks.
    ``total_scale``: samples the whole job processes per step (default: n_samples x world)."""
    torch, dist, dev, rank, world, lib = (ctx[k] for k in ("torch", "dist", "dev", "rank", "world", "lib"))
    from ignnition_b200 import Engine, ModelDescription
    from ignnition_b200.batching import assemble_tiled
    g, shape, qsize, _ = load_case(workload)
    dims = g["reference_meta"]["dimensions"]
    md = ModelDescription(g["model_json"], dims)
    eng = Engine(md, device=dev, seed=seed_weights,            # weights: the package's own seeded initialiser
                 fuse_sum_gru=(True if os.environ.get('IGN_FUSE_SUM_GRU') else None),
                 csr_mode=int(os.environ.get('IGN_CSR_MODE', '1')))
    base = g["reference_tensors"][0]
    out_entity = [o for o in md.get_readout_operations() if o.type == "predict"][0].input[0]
    batch = assemble_tiled(base, n_samples, eng.entities, eng.features, eng.adjacencies, eng.sequences,
                           feature_fns(qsize), seed=rank,
                           label_fn=(lambda r, n: r.normal(-1.0, 0.5, n)) if train else None,
                           label_entity=out_entity)
    pinned = eng.pack(batch)            # sample_of_* and the seq_* of destination-ordered lists stay on the host
    edges_per_iter = sum(batch.n_edges[a.name] for a in eng.adjacencies)
    n_pred = batch.num[out_entity]
    n_pred_glob = torch.tensor([n_pred], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_pred_glob)
    n_pred_glob = int(n_pred_glob.item())

    trainer = None
    if train:
        from ignnition_b200.train import Trainer
        trainer = Trainer(eng, world_size=world)

    def step_resident(graph):
        if graphed and trainer is not None:   # captured: build + forward + loss + backward; then all-reduce + optimiser
            return trainer.train_step_graphed(batch, pinned, global_n=n_pred_glob, copy=False)
        if graphed:                    # one captured CUDA graph: adjacency build + forward (Engine.forward_graphed)
            return eng.forward_graphed(batch, pinned, copy=False)
        if trainer is not None:        # model_fn train step: forward + loss + backward + all-reduce + Adam
            eng.build_graph(graph, training=True)
            graph.csr_t.clear()
            return trainer.train_step(graph, global_n=n_pred_glob)
        eng.build_graph(graph)
        return eng.forward(graph)

    host_pred = torch.empty(n_pred, 1, dtype=torch.float32, pin_memory=True)
    host_loss = torch.empty(4, dtype=torch.float64, pin_memory=True)

    # end to end: a stream of batches.  Batch k+1 is copied host -> device on a copy stream while batch k
    # is computed (two device staging buffers); every step still pays its own H2D and its own D2H.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)          # results go back while the next batch computes
    pred_done = torch.cuda.Event()
    stage = [torch.empty(pinned[0].numel(), dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"k": 0, "primed": False}

    def issue_upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[slot])          # the compute that last read this buffer is done
            g_ = eng.upload(batch, pinned, out=stage[slot])
            ready[slot].record(copy_stream)
        return g_

    def step_e2e_graphed():
        if trainer is not None:
            trainer.train_step_graphed(batch, pinned, global_n=n_pred_glob, copy=True)
            host_loss.copy_(trainer.scalars, non_blocking=True)
            return
        pred = eng.forward_graphed(batch, pinned, copy=True)       # H2D into the graph's staging buffer + one replay
        host_pred.copy_(pred, non_blocking=True)

    def step_e2e():
        if graphed:
            return step_e2e_graphed()
        k = e2e_state["k"]
        if not e2e_state["primed"]:
            for s_ in range(2):
                free[s_].record()
            e2e_state["next"] = issue_upload(k & 1)
            e2e_state["primed"] = True
        graph = e2e_state["next"]
        torch.cuda.current_stream().wait_event(ready[k & 1])
        e2e_state["next"] = issue_upload((k + 1) & 1)   # overlaps with this step's compute
        pred = step_resident(graph)
        free[k & 1].record()
        pred_done.record()
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(pred_done)
            if trainer is not None:                     # the step's result is the loss (sse, reg, count)
                host_loss.copy_(trainer.scalars, non_blocking=True)
            elif torch.is_tensor(pred) and pred.numel() == host_pred.numel():
                host_pred.copy_(pred, non_blocking=True)
                pred.record_stream(d2h_stream)
        e2e_state["k"] = k + 1
        return graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    graph = eng.upload(batch, pinned)
    for _ in range(max(warmup, 3)):
        step_resident(graph)
    barrier()
    sampler = ClockSampler(ctx["local"]) if (rank == 0 and ctx.get("clocks")) else None
    if sampler:
        sampler.start()
    l0 = lib.ign_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step_resident(graph)
    ev1.record()
    torch.cuda.synchronize()
    launches = lib.ign_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ev0.elapsed_time(ev1))

    ms_e2e = None
    if want_e2e:
        for _ in range(2):
            step_e2e()
        barrier()
        ev0.record()
        for _ in range(steps):
            step_e2e()
        d2h_stream.synchronize()                    # the last step's result is on the host before the clock stops
        ev1.record()
        torch.cuda.synchronize()
        ms_e2e = max_over_ranks(ev0.elapsed_time(ev1))

    # per-kernel CUDA-event timing of more passes (same stream, after the timed region)
    kern = profile_kernels(eng, graph, torch, steps, trainer, n_pred_glob) if (want_kernels and not graphed) else {}
    if graphed and trainer is not None:  # kernels inside the graph x replays + the optimiser's own launches (counted)
        launches += list(trainer._graphs.values())[0][4] * steps
    elif graphed:                       # launches were counted at capture time: kernels inside the graph x replays
        launches = eng.graphed_kernels(batch, pinned) * steps

    # parity of the timed configuration: the first samples of this rank's batch vs the fp64 CPU oracle, predictions
    # AND every entity's final state (north_star: 1e-5 on node and link states and predictions)
    parity = None
    if rank == 0 and not train and parity_samples:
        from oracle import ignnition_oracle as orc          # checker only
        eng.build_graph(graph)
        pred_t, states_t = eng.forward(graph, return_states=True)
        pred = pred_t.cpu().numpy().reshape(n_samples, -1)
        states = {e: states_t[e].cpu().numpy() for e in eng.entities}
        o64 = orc.Oracle(g["model_json"], dims, dtype=np.float64)
        o32 = orc.Oracle(g["model_json"], dims, dtype=np.float32)
        w32 = eng.get_weights()
        worst_p, worst_s = 0.0, {e: 0.0 for e in eng.entities}
        floor_p, floor_s = 0.0, {e: 0.0 for e in eng.entities}     # the NumPy fp32 oracle against the fp64 one
        k_chk = min(parity_samples, n_samples)
        for k in range(k_chk):
            t = dict(base)
            for name, ent, size in eng.features:
                n_e = int(base["num_" + ent])
                t[name] = batch.arrays["feat_" + name][k * n_e * size:(k + 1) * n_e * size]
            want, want_s = o64.forward(t, w32, return_states=True)
            want = want.reshape(-1)
            f32p, f32s = o32.forward(t, w32, return_states=True)
            worst_p = max(worst_p, float(np.abs(pred[k] - want).max() / np.abs(want).max()))
            floor_p = max(floor_p, float(np.abs(f32p.reshape(-1) - want).max() / np.abs(want).max()))
            for e in eng.entities:
                n_e = int(base["num_" + e])
                got = states[e][k * n_e:(k + 1) * n_e]
                worst_s[e] = max(worst_s[e], float(np.abs(got - want_s[e]).max() / np.abs(want_s[e]).max()))
                floor_s[e] = max(floor_s[e], float(np.abs(f32s[e] - want_s[e]).max() / np.abs(want_s[e]).max()))
        parity = {"max_rel_err_vs_fp64_oracle": worst_p, "state_max_rel_err_vs_fp64_oracle": worst_s,
                  "samples_checked": k_chk, "tolerance": 1e-5,
                  "norm": "max |a - b| / max |b| per sample and tensor",
                  "within_tolerance": bool(worst_p < 1e-5 and all(v < 1e-5 for v in worst_s.values())),
                  # what fp32 arithmetic itself costs on these inputs and weights: the reference's own program in
                  # NumPy fp32 against the same program in fp64 (profiles/r2_parity.md)
                  "fp32_oracle_vs_fp64_oracle": {"predictions": floor_p, "states": floor_s},
                  "weights": "Keras default initialisers (glorot kernels, orthogonal recurrent kernels), seed 0"}
        if graphed:
            parity["graph_replay_equals_eager_forward"] = bool(torch.equal(eng.forward_graphed(batch, pinned), pred_t))

    total = total_scale if total_scale is not None else n_samples * world
    out = {"workload": workload, "mode": ("train, one CUDA graph per step + optimiser update" if graphed and train else
                                          "inference, one CUDA graph per step" if graphed else
                                          "train" if train else "inference"), "n_gpus": world,
           "samples_per_gpu": n_samples, "samples_per_step": total, "steps": steps,
           "value": total * steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / steps,
           "mp_edges_per_s": edges_per_iter * eng.T * world * steps / (ms / 1e3),
           "gpu_launches": int(launches), "launches_per_step": launches / steps,
           "iterations": eng.T, "paths_per_gpu": batch.num.get("path"), "links_per_gpu": batch.num.get("link"),
           "mp_edges_per_iteration_per_gpu": edges_per_iter,
           "h2d_bytes_per_step": int(pinned[0].numel()),
           "d2h_bytes_per_step": int(host_loss.numel() * 8 if train else host_pred.numel() * 4),
           "kern": kern, "clocks": clocks, "parity": parity}
    if ms_e2e is not None:
        out["e2e"] = {"value": total * steps / (ms_e2e / 1e3), "unit": "samples/s",
                      "h2d_bytes_per_step": out["h2d_bytes_per_step"], "d2h_bytes_per_step": out["d2h_bytes_per_step"],
                      "ms_per_step": ms_e2e / steps}
    del graph, stage, eng, trainer
    torch.cuda.empty_cache()
    return out


def run_ingest_leg(ctx, n_per_file=256, files=16):
    """data.json TEXT -> predictions: the C++ ingest parses `files` dataset files of `n_per_file` GEANT2-shaped samples
    on host threads while the GPU runs the batches that are ready (upload + adjacency build + forward + D2H per file).
    Wall clock, rank 0's GPU; the text is built once outside the timed region."""
    import time as _t
    from concurrent.futures import ThreadPoolExecutor
    torch, dev = ctx["torch"], ctx["dev"]
    from ignnition_b200 import Engine, ModelDescription, synthetic
    from ignnition_b200.generator import sample_dimensions
    from ignnition_b200.ingest import NativeIngest
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "routenet_geant2.json")))
    samples = [synthetic.routenet_sample("geant2", s % 16, s) for s in range(n_per_file)]
    md = ModelDescription(g["model_json"], sample_dimensions(samples[0]))
    eng = Engine(md, device=dev, seed=0)
    text = json.dumps(samples).encode()
    workers = max(1, min(files, (os.cpu_count() or 1)))

    def parse(_):
        ing = NativeIngest(eng, None)
        ing.parse(text)
        return ing.batch()

    def run():
        n = 0
        with ThreadPoolExecutor(max_workers=workers) as pool:
            for b in pool.map(parse, range(files)):
                pred = eng.forward(eng.build_graph(eng.upload(b)))
                n += int(pred.numel())
                host = pred.cpu()
        torch.cuda.synchronize()
        return n

    run()
    t0 = _t.perf_counter()
    n_pred = run()
    dt = _t.perf_counter() - t0
    t1 = _t.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(parse, range(files)))
    dt_parse = _t.perf_counter() - t1
    total = n_per_file * files
    return {"workload": "ingest_text_to_predictions_geant2", "mode": "data.json text -> C++ ingest -> upload -> forward -> host",
            "samples": total, "json_mb": round(len(text) * files / 1e6, 1), "host_threads": workers,
            "host_cores": os.cpu_count(), "value": total / dt, "unit": "samples/s", "predictions": n_pred,
            "parse_only_samples_per_s": total / dt_parse,
            "note": "host-bound: the GPU forward of the same samples runs at the default line's rate"}


def load_json(path, default):
    try:
        return json.load(open(os.path.join(ROOT, path)))
    except Exception:
        return default


def roofline_of(case, hbm_peak, peak_src, default_size):
    """roofline object of the case's dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak"""
    kern = case["kern"]
    if not kern:
        return None
    top = max(kern.values(), key=lambda k: k["ms_total"])
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures of this workload at its
    # default size: profiles/ncu_traffic.json names the report every number comes from
    traffic = load_json("profiles/ncu_traffic.json", {}).get(case["workload"] + ("/train" if case["mode"] == "train" else ""), {})
    t = traffic.get(top["name"]) if default_size else None
    return {"bound": "hbm", "kernel": top["name"], "achieved": top["gbs"], "peak": hbm_peak, "unit": "GB/s",
            "frac": top["gbs"] / hbm_peak, "traffic": (t or {}).get("bytes_per_launch"),
            "traffic_source": (t or {}).get("source"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": top["bytes"], "avg_launch_ms": top["ms_avg"],
            "share_of_step": top["ms_total"] / max(sum(k["ms_total"] for k in kern.values()), 1e-9)}


def compact(case, hbm_peak, peak_src, scaling="weak", default_size=True, note=None):
    """an `also` entry: the same fields as the main line, without the per-kernel list"""
    out = {k: case[k] for k in ("workload", "mode", "n_gpus", "samples_per_gpu", "samples_per_step", "steps", "value",
                                "unit", "ms_per_step", "gpu_launches", "launches_per_step")}
    out["metric"] = "routenet_train_samples_per_s" if case["mode"] == "train" else "routenet_samples_per_s"
    out["top_kernels"] = []
    out["scaling"] = scaling
    if "e2e" in case:
        out["e2e"] = case["e2e"]
    out["roofline"] = roofline_of(case, hbm_peak, peak_src, default_size)
    top3 = sorted(case["kern"].values(), key=lambda k: -k["ms_total"])[:4]
    out["top_kernels"] = [{"name": k["name"], "launches_per_step": k["launches_per_step"],
                           "ms_total": round(k["ms_total"], 4), "gbs": round(k["gbs"], 1)} for k in top3]
    if case.get("parity"):
        out["parity"] = case["parity"]
    if note:
        out["note"] = note
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ignnition_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    ctx = {"torch": torch, "dist": dist, "dev": dev, "rank": rank, "world": world, "lib": lib, "local": local,
           "clocks": True}
    n_default = WORKLOADS[args.workload][3]
    n_samples = args.batch or n_default
    main = measure_case(ctx, args.workload, n_samples, args.train, args.steps, args.warmup, parity_samples=16)
    ctx["clocks"] = False

    peaks = load_json("MEASURED_PEAKS.json", {})
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"

    also = []
    if not args.no_also and not args.train:
        def leg(fn, name):
            try:
                also.append(fn())
            except Exception as exc:   # e.g. not enough free memory on a shared box: report, do not hide
                also.append({"workload": name, "error": str(exc)[:300]})
        # config 5: one big graph, destination-partitioned over the ranks (every rank takes part)
        leg(lambda: run_mpnn(args.mpnn_nodes, args.mpnn_edges, 64, 5, 3, torch, dev, "uniform", rank, world),
            "mpnn_uniform")
        if world > 1:      # SURVEY 8d variant C: 90 % of the edges stay inside the owner's rows, boundary rows only
            leg(lambda: run_mpnn(args.mpnn_nodes, args.mpnn_edges, 64, 5, 3, torch, dev, "local", rank, world,
                                 "boundary"), "mpnn_local")
        if rank == 0:      # from dataset text to predictions, host ingest + GPU
            leg(lambda: run_ingest_leg(ctx), "ingest_text_to_predictions")
        # config 4: RouteNet synth50 training, gradient all-reduce inside the step (weak scaling: 256 samples per GPU)
        leg(lambda: compact(measure_case(ctx, "routenet_synth50_b256", 256, True, 10, 3), hbm_peak, peak_src),
            "routenet_synth50_b256/train")
        # config 3 training at the headline size (the BPTT kernels): 4096 samples per GPU
        leg(lambda: compact(measure_case(ctx, "routenet_geant2_b4096", 4096, True, 6, 3), hbm_peak, peak_src),
            "routenet_geant2_b4096/train")
        # config 3 as BASELINE words it: batch 4096 data-parallel OVER the GPUs (strong scaling)
        if world > 1:
            leg(lambda: compact(measure_case(ctx, "routenet_geant2_b4096", 4096 // world, False, args.steps, 3,
                                             total_scale=(4096 // world) * world),
                                hbm_peak, peak_src, scaling="strong", default_size=False,
                                note="4096 samples split over the GPUs, no collective"),
                "routenet_geant2_b4096/strong")
            leg(lambda: compact(measure_case(ctx, "routenet_geant2_b4096", 4096 // world, False, args.steps, 3,
                                             total_scale=(4096 // world) * world, graphed=True),
                                hbm_peak, peak_src, scaling="strong", default_size=False,
                                note="4096 samples split over the GPUs, no collective; the step replayed as one captured "
                                     "CUDA graph (Engine.forward_graphed)"),
                "routenet_geant2_b4096/strong/graphed")
        # config 2: Q-size RouteNet (links + paths + nodes, interleave aggregation)
        leg(lambda: compact(measure_case(ctx, "qsize_nsfnet_b4096", 4096, False, 10, 3, parity_samples=16), hbm_peak,
                            peak_src), "qsize_nsfnet_b4096")
        # config 1: RouteNet on the NSFNET sample at the reference's own batch sizes (train_options.ini: 3 / 32)
        for bsz in (3, 32):
            for tr in (False, True):
                leg(lambda: compact(measure_case(ctx, "routenet_nsfnet_b4096", bsz, tr, 30, 5, parity_samples=0 if tr else 3),
                                    hbm_peak, peak_src, default_size=False,
                                    note="launch-bound size: %d samples per step" % bsz),
                    "routenet_nsfnet_b%d/%s" % (bsz, "train" if tr else "inference"))
            leg(lambda: compact(measure_case(ctx, "routenet_nsfnet_b4096", bsz, False, 100, 5, parity_samples=3,
                                             graphed=True), hbm_peak, peak_src, default_size=False,
                                note="%d samples per step; adjacency build + T = 8 iterations + readout replayed as one "
                                     "captured CUDA graph (Engine.forward_graphed)" % bsz),
                "routenet_nsfnet_b%d/graphed" % bsz)
            leg(lambda: compact(measure_case(ctx, "routenet_nsfnet_b4096", bsz, True, 100, 5, graphed=True), hbm_peak,
                                peak_src, default_size=False,
                                note="%d samples per train step; adjacency build + forward + loss + backward replayed as one "
                                     "captured CUDA graph, optimiser update after it (Trainer.train_step_graphed)" % bsz),
                "routenet_nsfnet_b%d/train/graphed" % bsz)

    if rank == 0:
        roof = roofline_of(main, hbm_peak, peak_src, n_samples == n_default)
        kern = main["kern"]
        top = max(kern.values(), key=lambda k: k["ms_total"])
        cores = os.cpu_count() or 1
        cpu_val, cpu_n, cpu_dt = cpu_samples_per_s(args.workload, args.cpu_samples, 1)
        line = {
            "metric": "routenet_train_samples_per_s" if args.train else "routenet_samples_per_s",
            "value": main["value"],
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "samples_per_gpu": n_samples, "iterations": main["iterations"],
                       "paths_per_gpu": main["paths_per_gpu"], "links_per_gpu": main["links_per_gpu"],
                       "mp_edges_per_iteration_per_gpu": main["mp_edges_per_iteration_per_gpu"],
                       "timing": "inputs+states per step exceed L2 (no flush needed)",
                       "step": ("device CSR build (+ transposed) + forward + MSE/l2 + backward + NCCL all-reduce + Adam"
                                if args.train else "device CSR build + T message-passing iterations + readout")},
            "mp_edges_per_s": main["mp_edges_per_s"],
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "roofline": roof,
            "kernels": sorted(kern.values(), key=lambda k: -k["ms_total"]),
            "cpu_baseline": {"value": cpu_val, "unit": "samples/s", "cores": 1, "kind": "port",
                             "host_cores": cores,
                             "sample": "%d samples of %s through oracle/ignnition_oracle.py (NumPy fp32, "
                                       "per-sample loop), %.1f s" % (cpu_n, args.workload, cpu_dt)},
            "clocks": main["clocks"],
            "parity": main["parity"],
            "also": also or None,
        }
        if top["name"] == "ign_gru_seq" and not args.train and args.workload.startswith("routenet"):
            # the ordered update is a chain of small GEMMs: the tensor-pipe view of the same launch.  FLOPs issued =
            # 3 (3xTF32 products) x 2 E 3U (F + U); peak = the tcgen05 kind::tf32 rate MEASURED on this chip
            # (profiles/measured_tf32_peak.json, tools/umma_rate.cu), not a fraction of the bf16 number
            e_steps = main["mp_edges_per_iteration_per_gpu"] // 2    # path-link incidences = steps of one ordered update
            fl = 3.0 * 2.0 * e_steps * 96 * 64
            tf32 = load_json("profiles/measured_tf32_peak.json", {})
            tf32_peak = float(tf32.get("tf32_tflops", float(peaks.get("bf16_tflops", 1615.9)) / 2.0))
            ach = fl / (top["ms_avg"] * 1e-3) / 1e12
            line["roofline_tensor"] = {
                "bound": "tensor", "kernel": "ign_gru_seq", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": ach / tf32_peak, "flops_issued_per_launch": fl,
                "peak_source": "profiles/measured_tf32_peak.json" if tf32 else "MEASURED_PEAKS bf16_tflops / 2",
                "note": "tf32 operations issued (3 per fp32 product); the kernel is bound by the per-tile dependency "
                        "chain (profiles/r1_walk_phases.md)"}
        mp = also[0] if also else None
        if mp and "unfused_pair" in mp:
            line["roofline_gather_segment"] = {
                "bound": "hbm", "kernel": "ign_agg_gru_cell_tc (config 5: gather + segment sum + GRU, F = U = 64)",
                "achieved": mp["fused_update"]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": mp["fused_update"]["achieved_gbs"] / hbm_peak,
                "segment_reduce_alone": {"achieved": mp["unfused_pair"]["segment_reduce_gbs"],
                                         "frac": mp["unfused_pair"]["segment_reduce_gbs"] / hbm_peak},
                "traffic": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def profile_kernels(eng, graph, torch, reps, trainer=None, n_glob=None):
    """CUDA-event time of every launch class in one forward (events on the launching stream)."""
    from ignnition_b200 import ops
    records = {}
    orig = {}

    def wrap(name, bytes_fn):
        fn = getattr(ops, name)
        orig[name] = fn

        def timed(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            records.setdefault(name, []).append((e0, e1, bytes_fn(*a, **kw)))
            return out
        setattr(ops, name, timed)

    def b_gru_seq(steps_rowptr, steps, order, srcs, h0, *a, **kw):
        E, n, F, U = steps.numel(), h0.shape[0], srcs[0].shape[1], h0.shape[1]
        return E * (4 + 4 * F) + n * (8 * U + 8)            # col + gathered row per step; h0 in, h out, rowptr, order

    def b_agg(rowptr, col, src_states, h_dst, *a, **kw):
        E, n, F, U = col.numel(), h_dst.shape[0], src_states.shape[1], h_dst.shape[1]
        return E * (4 + 4 * F) + n * (8 * U + 4)

    def b_agg_tc(op, rowptr, col, src_states, h_dst, *a, **kw):
        E, n, F, U = col.numel(), h_dst.shape[0], src_states.shape[1], h_dst.shape[1]
        return E * (4 + 4 * F) + n * (8 * U + 4)             # col + gathered row per slot; old state in, new state out, rowptr

    def b_seg(op, rowptr, col, src_states, *a, **kw):
        n, F = rowptr.numel() - 1, src_states.shape[1]
        E = col.numel() if col is not None else src_states.shape[0]
        return E * (4 + 4 * F) + n * (4 * F + 4)

    def b_dense(x, w, bias, act, *a, **kw):
        return 4 * (x.numel() + x.shape[0] * w.shape[1] + w.numel())

    def b_csr(dst, src, seq, num_dst, *a, **kw):
        return dst.numel() * 8 * 2 * 3 + num_dst * 4

    def b_gru_cell(x, h, *a, **kw):
        return 4 * (x.numel() + 2 * h.numel())

    zero = lambda *a, **kw: 0

    def b_seq_bwd_steps(plan, meta, max_steps, srcs, h0, h_seq, *a, **kw):
        # per step: x, h, dL/dh gathered, dx and dL/dh written (5 rows), G (4 rows) written and read once, [x | h]
        # read again by the weight-gradient kernel
        E, F = h_seq.shape[0], h0.shape[1]
        return E * (5 * 4 * F + 2 * 16 * F + 8 * F)

    def b_seq_bwd(steps_rowptr, steps, order, srcs, h0, h_seq, *a, **kw):
        E, F = steps.numel(), h0.shape[1]
        return E * 5 * 4 * F

    def b_dense_bwd(x, w, act, pre_act, dy, dx, dw, db):
        m, k = x.shape
        n = w.shape[1]                       # dZ in place (r + w, pre-activation read), dX written, X and dZ read for dW
        return 4 * m * ((3 * n if pre_act is not None else 0) + n + (k if dx is not None else 0) + k + n)

    def b_gru_cell_bwd(x, h, *a, **kw):
        return 4 * (2 * x.numel() + 3 * h.numel())

    def b_mlp_head(x, w1, b1, act1, w2, b2, act2, w3, b3, *a, **kw):
        return 4 * (x.numel() + x.shape[0] + w1.numel() + w2.numel())      # x in, one float per row out

    def b_dense_head(x, w, bias, act, head_w, head_b, *a, **kw):
        return 4 * (x.numel() + x.shape[0] + w.numel())

    def b_init(feats, sizes, n, hidden, *a, **kw):
        return 4 * (n * hidden + sum(int(f.numel()) for f in feats))

    for name, fn in (("gru_seq", b_gru_seq), ("agg_gru_cell", b_agg), ("agg_gru_cell_tc", b_agg_tc), ("segment_reduce", b_seg),
                     ("dense", b_dense), ("csr_build", b_csr), ("gru_cell", b_gru_cell),
                     ("mlp_head", b_mlp_head), ("dense_head", b_dense_head), ("init_state", b_init),
                     ("length_order", zero), ("seq_meta", zero), ("steps_build", zero),
                     ("gru_seq_steps", zero), ("seq_step_plan", zero),
                     ("gru_seq_bwd", b_seq_bwd), ("gru_seq_bwd_steps", b_seq_bwd_steps),
                     ("gru_cell_bwd", b_gru_cell_bwd), ("dense_bwd", b_dense_bwd),
                     ("dense_head_bwd_chain", lambda x, *a, **kw: 8 * x.numel())):
        wrap(name, fn)
    try:
        for _ in range(max(1, min(reps, 3))):
            if trainer is not None:
                eng.build_graph(graph, training=True)
                graph.csr_t.clear()
                trainer.train_step(graph, global_n=n_glob)
            else:
                eng.build_graph(graph)
                eng.forward(graph)
        torch.cuda.synchronize()
    finally:
        for name, fn in orig.items():
            setattr(ops, name, fn)
    out = {}
    n_pass = max(1, min(reps, 3))
    for name, recs in records.items():
        ms = [a.elapsed_time(b) for a, b, _ in recs]
        by = [c for _, _, c in recs]
        tot = float(sum(ms))
        out[name] = {"name": "ign_" + name, "launches_per_step": len(recs) // n_pass, "ms_total": tot / n_pass,
                     "ms_avg": tot / len(recs), "bytes": float(np.mean(by)),
                     "gbs": float(sum(by)) / (tot / 1e3) / 1e9 if tot > 0 else 0.0}
    return out


# ------------------------------------------------------------------------------ config 5: big graph
def mpnn_model_json(hidden=64, iterations=8):
    """generic sum-aggregation MPNN (BASELINE config 5) in the reference's JSON keywords"""
    return {
        "entities": [{"name": "node", "hidden_state_dimension": hidden,
                      "features": [{"name": "x", "normalization": "None"}]}],
        "message_passing": {"num_iterations": iterations, "stages": [{"stage_name": "s", "stage_mp": [{
            "destination_entity": "node",
            "source_entities": [{"name": "node", "adj_vector": "adj", "message": [{"type": "direct_assignation"}]}],
            "aggregation": {"type": "sum"},
            "update": {"type": "recurrent_neural_network", "nn_name": "rec"}}]}]},
        "readout": [{"type": "predict", "input": ["node"], "label": "y", "nn_name": "ro"}],
        "neural_networks": [
            {"nn_name": "rec", "nn_type": "recurrent_neural_network", "recurrent_type": "GRU"},
            {"nn_name": "ro", "nn_type": "feed_forward", "nn_architecture": [
                {"type_layer": "Dense", "units": 1, "activation": "None"}]}],
        "learning_options": {"loss": "MeanSquaredError", "optimizer": {"type": "Adam"}},
    }


def mpnn_shard(n_nodes, n_edges, hidden, variant, rank, world, torch, dev, seed=0):
    """This rank's contiguous shard of ONE global synthetic graph (the same graph for every world size):
    the edge list and the node features are generated in fixed chunks, each seeded by its global chunk
    index, so rank r of W produces exactly the pieces [r/W, (r+1)/W) of the list a single rank produces."""
    e_chunk, n_chunk = 1_000_000, 50_000
    while n_edges % (e_chunk * world) and e_chunk > 1:
        e_chunk //= 2
    while n_nodes % (n_chunk * world) and n_chunk > 1:
        n_chunk //= 2
    gen = torch.Generator(device=dev)
    srcs, dsts = [], []
    for c in range(rank * (n_edges // world) // e_chunk, (rank + 1) * (n_edges // world) // e_chunk):
        gen.manual_seed(seed * 1_000_003 + 2 * c)
        src = torch.randint(0, n_nodes, (e_chunk,), device=dev, dtype=torch.int32, generator=gen)
        if variant == "skew":                # power-law-like in-degrees: dst = floor(N u^3)
            u = torch.rand(e_chunk, device=dev, generator=gen)
            dst = (u * u * u * n_nodes).to(torch.int32).clamp_(0, n_nodes - 1)
        else:
            dst = torch.randint(0, n_nodes, (e_chunk,), device=dev, dtype=torch.int32, generator=gen)
        if variant == "local":               # SURVEY 8d variant C: 90 % of the sources live in the destination's
            blk = n_nodes // 8               # eighth of the node range (the owner's rows at 8 GPUs)
            loc = torch.rand(e_chunk, device=dev, generator=gen) < 0.9
            near = (dst // blk) * blk + torch.randint(0, blk, (e_chunk,), device=dev, dtype=torch.int32, generator=gen)
            src = torch.where(loc, near.clamp_(0, n_nodes - 1), src)
        srcs.append(src)
        dsts.append(dst)
    feats = []
    for c in range(rank * (n_nodes // world) // n_chunk, (rank + 1) * (n_nodes // world) // n_chunk):
        gen.manual_seed(seed * 1_000_003 + 2 * c + 1)
        feats.append(torch.randn(n_chunk, hidden, device=dev, generator=gen))
    return torch.cat(srcs), torch.cat(dsts), torch.cat(feats)


def run_mpnn(n_nodes, n_edges, hidden, steps, warmup, torch, dev, variant="uniform", rank=0, world=1,
             exchange="copy"):
    """Message-passing iterations of the generic MPNN on ONE large synthetic graph (strong scaling) through
    the product path ``ignnition_b200.parallel.PartitionedEngine``: rank r owns a contiguous range of
    destination rows; one fused kernel per iteration gathers, sums, applies the GRU and stores the new rows
    into every rank's peer-mapped state array (exchange 'peer'), or hands them to ncclAllGather ('nccl') or
    sends only the rows a peer reads ('boundary').  Returns edges/s per iteration of the WHOLE graph (max
    over ranks) with the HBM roofline of the kernel and the NVLink floor of the exchange."""
    import torch.distributed as dist
    from ignnition_b200 import Engine, ModelDescription, ops
    from ignnition_b200.parallel import PartitionedEngine
    n_nodes = n_nodes // (8 * 128) * (8 * 128)
    n_edges = n_edges // 8 * 8
    md = ModelDescription(mpnn_model_json(hidden), {"x": hidden, "adj": 0})
    eng = Engine(md, device=dev, seed=0)
    src, dst, x = mpnn_shard(n_nodes, n_edges, hidden, variant, rank, world, torch, dev)
    pe = PartitionedEngine(eng, exchange=exchange if world > 1 else "copy")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pe.build({"node": n_nodes}, {"adj": (src, dst)}, {"x": x})
    e1.record()
    torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    del src, dst, x
    own = pe.own("node")[1] - pe.own("node")[0]
    e_own = pe.n_edges["adj"]
    for _ in range(max(warmup, 3)):
        pe.message_passing(iterations=1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    pe.exchanged_bytes = 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pe.message_passing(iterations=steps)
    t1.record()
    torch.cuda.synchronize()
    t = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    recv = pe.exchanged_bytes / steps
    chk = pe.checksum()["node"]
    # the fused kernel alone on this rank's rows (no exchange): its HBM roofline
    rowptr, col = pe.csr["adj"]
    lo, hi = pe.own("node")
    K, R, B = (eng.param("node_update/kernel"), eng.param("node_update/recurrent_kernel"), eng.param("node_update/bias"))
    scratch = torch.empty(hi, hidden, device=dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(steps, 3)
    for _ in range(2):
        ops.agg_gru_cell_tc(ops.OP_SUM, rowptr, col, pe.full_state("node"), pe.state("node"), K, R, B, [scratch], out_row0=lo)
    k0.record()
    for _ in range(reps):
        ops.agg_gru_cell_tc(ops.OP_SUM, rowptr, col, pe.full_state("node"), pe.state("node"), K, R, B, [scratch], out_row0=lo)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / reps
    # the round-1 pair on the same rows, for the ablation: gather + segment sum, then the GRU cell
    agg = torch.empty(own, hidden, device=dev)
    p0, p1, p2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ops.segment_reduce(ops.OP_SUM, rowptr, col, pe.full_state("node"), out=agg)
    ops.gru_cell(agg, pe.state("node"), K, R, B, out=scratch[lo:hi])
    p0.record()
    for _ in range(reps):
        ops.segment_reduce(ops.OP_SUM, rowptr, col, pe.full_state("node"), out=agg)
    p1.record()
    for _ in range(reps):
        ops.gru_cell(agg, pe.state("node"), K, R, B, out=scratch[lo:hi])
    p2.record()
    torch.cuda.synchronize()
    seg_ms, cell_ms = p0.elapsed_time(p1) / reps, p1.elapsed_time(p2) / reps
    del scratch, agg
    kern_bytes = e_own * (4 + 4 * hidden) + own * (8 * hidden + 4)
    seg_bytes = e_own * (4 + 4 * hidden) + own * (4 * hidden + 4)
    out = {"workload": "mpnn_%s_n%d_e%d_h%d" % (variant, n_nodes, n_edges, hidden), "n_gpus": world,
           "path": "ignnition_b200.parallel.PartitionedEngine", "exchange": pe.exchange if world > 1 else "none",
           "partition": "destination-node range, %d nodes / %d in-edges on rank 0" % (own, e_own),
           "mp_edges_per_s_per_iteration": n_edges * steps / (total_ms / 1e3),
           "ms_per_iteration": total_ms / steps, "iterations_timed": steps,
           "build_ms": build_ms, "build_edges_per_s": n_edges / (build_ms / 1e3),
           "fused_update": {"kernel": "ign_agg_gru_cell_tc (gather + sum + GRU + TMA stores)", "avg_launch_ms": kern_ms,
                            "algorithmic_bytes_per_launch": kern_bytes,
                            "achieved_gbs": kern_bytes / (kern_ms / 1e3) / 1e9},
           "unfused_pair": {"segment_reduce_ms": seg_ms, "gru_cell_ms": cell_ms,
                            "segment_reduce_gbs": seg_bytes / (seg_ms / 1e3) / 1e9},
           "state_checksum": chk}
    if world > 1:
        floor_ms = recv / 770e9 * 1e3
        out["exchange_detail"] = {"bytes_received_per_gpu_per_iteration": recv, "nvlink_peer_peak_gbs": 770.0,
                                  "nvlink_floor_ms": floor_ms,
                                  "floor_over_achieved": max(floor_ms, kern_ms) / (total_ms / steps)}
    pe.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--cpu-samples", type=int, default=48, help="samples of the bounded CPU-baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the config-5 big-graph leg of the default run")
    ap.add_argument("--train", action="store_true",
                    help="time the train step (forward + loss + backward + gradient all-reduce + Adam) instead of inference")
    ap.add_argument("--mpnn-nodes", type=int, default=10_000_000)
    ap.add_argument("--mpnn-edges", type=int, default=200_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
