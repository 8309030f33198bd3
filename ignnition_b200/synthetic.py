"""Seeded synthetic datasets in IGNNITION's on-disk format.

The reference's example datasets are not in its tree (``.MISSING_LARGE_BLOBS``), so the BASELINE
configs are synthesised: NSFNET-/GEANT2-/synth50-shaped RouteNet samples and Q-size samples in the
exact ``data.json`` layout ``examples/Routenet/migrate.py:54-109`` (``process_sample``) writes:
``entities`` (links first, then paths), ``adj_links_paths`` (path -> ordered links),
``adj_paths_links`` (link -> paths, keys in order of first use), per-path ``traffic`` / ``delay`` /
``jitter`` and per-link ``link_capacity``.
"""

from __future__ import annotations

import io
import json
import os
import tarfile
from collections import deque
from typing import Dict, List, Sequence, Tuple

import numpy as np

SHAPES = {
    # name: (nodes, undirected edges)
    "nsfnet": (14, 21),
    "geant2": (24, 37),
    "synth50": (50, 138),
}


def random_topology(n_nodes: int, n_undirected: int, seed: int) -> List[Tuple[int, int]]:
    """Connected G(n, m): random spanning tree plus random extra edges; returns directed links."""
    rng = np.random.RandomState(seed)
    order = rng.permutation(n_nodes)
    und = set()
    for k in range(1, n_nodes):
        a = int(order[k]); b = int(order[rng.randint(0, k)])
        und.add((min(a, b), max(a, b)))
    while len(und) < n_undirected:
        a, b = rng.randint(0, n_nodes, size=2)
        if a != b:
            und.add((int(min(a, b)), int(max(a, b))))
    links = []
    for a, b in sorted(und):
        links.append((a, b))
        links.append((b, a))
    return links


def shortest_path_routing(n_nodes: int, links: Sequence[Tuple[int, int]], seed: int = 0):
    """BFS shortest path (node list) for every ordered pair; neighbour order shuffled by seed."""
    rng = np.random.RandomState(seed)
    nbrs: List[List[int]] = [[] for _ in range(n_nodes)]
    for a, b in links:
        nbrs[a].append(b)
    for l in nbrs:
        rng.shuffle(l)
    routing: Dict[Tuple[int, int], List[int]] = {}
    for s in range(n_nodes):
        prev = [-1] * n_nodes
        prev[s] = s
        dq = deque([s])
        while dq:
            u = dq.popleft()
            for v in nbrs[u]:
                if prev[v] < 0:
                    prev[v] = u
                    dq.append(v)
        for t in range(n_nodes):
            if t == s:
                continue
            path = [t]
            while path[-1] != s:
                path.append(prev[path[-1]])
            routing[(s, t)] = path[::-1]
    return routing


def routenet_sample(shape: str = "nsfnet", topo_seed: int = 0, feat_seed: int = 0,
                    qsize: bool = False) -> dict:
    """One sample dict in the reference's data.json format.

    ``qsize=True`` adds the ``node`` entity of ``examples/Q-size/model_description.json``: the
    egress queue at the source node of every hop, ``adj_nodes_paths`` / ``adj_paths_nodes``,
    ``queue_sizes`` and ``path_interleave = ["node", "link"]``.
    """
    n_nodes, n_und = SHAPES[shape]
    links = random_topology(n_nodes, n_und, topo_seed)
    routing = shortest_path_routing(n_nodes, links, topo_seed)
    rng = np.random.RandomState(feat_seed)
    n_paths = n_nodes * (n_nodes - 1)
    data: dict = {}
    if qsize:
        data["traffic"] = rng.uniform(0.05, 0.6, n_paths).tolist()
        data["delay"] = rng.lognormal(-1.78, 0.93, n_paths).tolist()
        data["jitter"] = rng.uniform(0.1, 3.0, n_paths).tolist()
        data["link_capacity"] = rng.choice([10.0, 25.0, 40.0], len(links)).tolist()
        data["queue_sizes"] = rng.choice([1.0, 8.0, 16.0, 32.0], n_nodes).tolist()
    else:
        data["traffic"] = rng.uniform(40.0, 300.0, n_paths).tolist()
        data["delay"] = rng.lognormal(-1.0, 0.5, n_paths).tolist()
        data["jitter"] = rng.uniform(0.01, 1.0, n_paths).tolist()
        data["link_capacity"] = rng.choice([10000.0, 40000.0], len(links)).tolist()
    ent: Dict[str, str] = {}
    link_name = {}
    for i, (a, b) in enumerate(links):
        ent["l%d" % i] = "link"
        link_name[(a, b)] = "l%d" % i
    if qsize:
        for v in range(n_nodes):
            ent["n%d" % v] = "node"
    data["entities"] = ent
    apl: Dict[str, list] = {}      # link -> paths
    alp: Dict[str, list] = {}      # path -> links
    anp: Dict[str, list] = {}      # path -> nodes
    apn: Dict[str, list] = {}      # node -> paths
    p = 0
    for i in range(n_nodes):
        for j in range(n_nodes):
            if i == j:
                continue
            pn = "p%d" % p
            ent[pn] = "path"
            nodes = routing[(i, j)]
            for a, b in zip(nodes[:-1], nodes[1:]):
                ln = link_name[(a, b)]
                apl.setdefault(ln, []).append(pn)
                alp.setdefault(pn, []).append(ln)
                if qsize:
                    nn = "n%d" % a
                    anp.setdefault(pn, []).append(nn)
                    apn.setdefault(nn, []).append(pn)
            p += 1
    data["adj_paths_links"] = apl
    data["adj_links_paths"] = alp
    if qsize:
        data["adj_nodes_paths"] = anp
        data["adj_paths_nodes"] = apn
        data["path_interleave"] = ["node", "link"]
    return data


def write_dataset(directory: str, samples: Sequence[dict], per_file: int = 25) -> List[str]:
    """``sample_<k>.tar.gz`` files each holding one ``data.json`` (list of samples)."""
    os.makedirs(directory, exist_ok=True)
    out = []
    for k in range(0, len(samples), per_file):
        payload = json.dumps(list(samples[k:k + per_file])).encode()
        path = os.path.join(directory, "sample_%d.tar.gz" % (k // per_file))
        with tarfile.open(path, "w:gz") as tar:
            info = tarfile.TarInfo("data.json")
            info.size = len(payload)
            tar.addfile(info, io.BytesIO(payload))
        out.append(path)
    return out
