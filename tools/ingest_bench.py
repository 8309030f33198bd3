"""Host-side ingest throughput: data.json text -> batch arrays.

python path : json.loads + generator.sample_to_tensors (mirror of the reference generator,
              generator_std_to_framework.py:53-230) + batching.assemble
native path : ignnition_b200.ingest.NativeIngest (csrc/ingest.cpp), one thread
Both from the same JSON text of N GEANT2-shaped RouteNet samples; the arrays are compared.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ignnition_b200 import synthetic                                   # noqa: E402
from ignnition_b200.batching import assemble                           # noqa: E402
from ignnition_b200.generator import sample_dimensions, sample_to_tensors   # noqa: E402
from ignnition_b200.ingest import NativeIngest                         # noqa: E402
from ignnition_b200.model_description import ModelDescription          # noqa: E402
from test_host import _SpecEngine                                      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mj = json.load(open("tests/golden/routenet_geant2.json"))["model_json"]
samples = [synthetic.routenet_sample("geant2", s, s) for s in range(n)]
md = ModelDescription(mj, sample_dimensions(samples[0]))
eng = _SpecEngine(md)
text = json.dumps(samples).encode()
out_name = md.get_output_info()[0]
feats = [f[0] for f in eng.features]

t0 = time.perf_counter()
parsed = json.loads(text)
t1 = time.perf_counter()
pairs = [sample_to_tensors(s, feats, out_name, md.get_adjecency_info(), [], [], True) for s in parsed]
t2 = time.perf_counter()
want = assemble([p[0] for p in pairs], eng.entities, eng.features, eng.adjacencies, (), [p[1] for p in pairs])
t3 = time.perf_counter()

import os
from concurrent.futures import ThreadPoolExecutor
py = t3 - t0
native = {}
for workers in sorted({1, 4, min(16, os.cpu_count() or 1)}):
    # `workers` handles, each parsing the same text concurrently (= that many dataset files at once)
    ings = [NativeIngest(eng, label_name=out_name) for _ in range(workers)]

    def one(ing):
        ing.reset()
        ing.parse(text)
        return ing.batch()

    best = 1e9
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for _ in range(3):
            a = time.perf_counter()
            got = list(pool.map(one, ings))
            best = min(best, time.perf_counter() - a)
    for k in want.arrays:
        assert np.array_equal(got[-1].arrays[k], want.arrays[k]), k
    native[workers] = {"samples_per_s": round(workers * n / best, 1), "mb_per_s": round(workers * len(text) / 1e6 / best, 1),
                       "speedup_vs_python_1_core": round(workers * py / best, 1)}
print(json.dumps({
    "samples": n, "json_mb": round(len(text) / 1e6, 2), "edges": int(sum(want.n_edges.values())),
    "python": {"samples_per_s": round(n / py, 1), "json_loads_s": round(t1 - t0, 3), "per_edge_loops_s": round(t2 - t1, 3),
               "assemble_s": round(t3 - t2, 3)},
    "native_by_concurrent_files": native, "host_cores": os.cpu_count(), "arrays_identical": True}))
