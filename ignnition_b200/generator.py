"""Host half of the adjacency builder: dataset sample -> index tensors.

Mirrors the contract of the reference's ``generator`` (``code/utils/generator_std_to_framework.py``
:53-230) and ``make_indices`` (:32-50): for every sample of a ``data.json`` it yields the dict

    <feature>                      list/array of floats          (:102-107)
    src_<adj>, dst_<adj>           int64 [E]                      (:174-178)
    seq_<src entity>_<dst entity>  int64 [E]  position of the edge inside its destination (:153, :181)
    params_<adj>                   [E, k] edge parameters when present (:156-163, :185)
    num_<entity>                   int                            (:188-190)
    indices_<entity>_to_<dst>      int64, interleave positions    (:193-219)

The arrays are the *input* of the device CSR builder (``csrc/csr_build.cu``); the on-disk format
(``*.tar.gz`` holding ``data.json``, ``examples/Routenet/migrate.py:54-109``) is unchanged.
"""

from __future__ import annotations

import glob
import json
import math
import random
import tarfile
from typing import Optional, Dict, Iterable, Iterator, List, Sequence, Tuple

import numpy as np


class DatasetError(Exception):
    pass


def make_indices(entities: Dict[str, str]) -> Tuple[Dict[str, int], Dict[str, int]]:
    """Entity-name -> dense index per entity type, in order of first appearance.

    Same result as the reference ``make_indices`` (generator_std_to_framework.py:32-50):
    returns ``(count per entity type, index per node name)``.
    """
    counter: Dict[str, int] = {}
    indices: Dict[str, int] = {}
    for node, kind in entities.items():
        n = counter.get(kind, 0)
        indices[node] = n
        counter[kind] = n + 1
    return counter, indices


def _adjacency(sample: dict, entities: Dict[str, str], indices: Dict[str, int], name: str,
               src_entity: str, dst_entity: str, uses_parameters: str):
    if name not in sample:
        raise DatasetError('A list for the adjecency vector named "' + name +
                           '" was not found although being expected.')
    src_idx: List[int] = []
    dst_idx: List[int] = []
    seq: List[int] = []
    params: list = []
    # grouped by destination in JSON insertion order (generator_std_to_framework.py:144-146)
    for destination, sources in sample[name].items():
        if entities[destination] != dst_entity:
            raise DatasetError(
                'The adjecency list "' + name + '" was expected to be from ' + src_entity + ' to ' +
                dst_entity + '.\n However, "' + destination + '" was found which is of type "' +
                entities[destination] + '" instead of ' + dst_entity)
        n = len(sources)
        d = indices[destination]
        seq.extend(range(n))
        dst_idx.extend([d] * n)
        if n and isinstance(sources[0], list):          # [[name, params], ...]
            src_idx.extend(indices[s[0]] for s in sources)
            if uses_parameters == 'True':
                params.extend(s[1] for s in sources)
        else:
            for s in sources:
                if entities[s] != src_entity:
                    raise DatasetError(
                        'The adjecency list "' + name + '" was expected to be from "' + src_entity +
                        '" to "' + dst_entity + '.\n However, "' + s + '" was found which is of type "' +
                        entities[s] + '" instead of "' + src_entity)
            src_idx.extend(indices[s] for s in sources)
    return src_idx, dst_idx, seq, params


def interleave_indices(pattern: Sequence[str], max_len: Dict[str, int]) -> Dict[str, List[int]]:
    """Positions of each entity in the tiled interleave pattern (generator :193-219).

    ``max_len[entity]`` is ``max(seq_<entity>_<dst>) + 1`` of that sample.
    """
    ids: Dict[str, int] = {}
    numeric: List[int] = []
    n_total = 0
    for ent in pattern:
        if ent not in ids:
            ids[ent] = len(ids)
            n_total += max_len[ent]
        numeric.append(ids[ent])
    reps = math.ceil(float(n_total) / len(pattern))
    tiled = np.array((numeric * reps)[:n_total])
    return {ent: np.where(tiled == i)[0].tolist() for ent, i in ids.items()}


def sample_to_tensors(sample: dict, feature_names: Iterable[str], output_name, adj_names,
                      interleave_names, additional_input, training: bool):
    """One dataset sample -> the reference's tensor dict (lists of python numbers)."""
    data: dict = {}
    for f in feature_names:
        if f not in sample:
            raise DatasetError('A list for feature named "' + str(f) +
                               '" was not found although being expected.')
        data[f] = sample[f]
    for a in additional_input:
        if a not in sample:
            raise DatasetError('The input name "' + str(a) + '" was not found although being expected.')
        data[a] = sample[a]
    output: list = []
    if training:
        if output_name not in sample:
            raise DatasetError('A list for the output named "' + str(output_name) +
                               '" was not found although being expected.')
        v = sample[output_name]
        output += v if isinstance(v, list) else [v]

    entities = sample['entities']
    counts, indices = make_indices(entities)
    seqs: Dict[str, List[int]] = {}
    for name, src_entity, dst_entity, uses_parameters in adj_names:
        s, d, q, p = _adjacency(sample, entities, indices, name, src_entity, dst_entity,
                                uses_parameters)
        data['src_' + name] = s
        data['dst_' + name] = d
        data['seq_' + src_entity + '_' + dst_entity] = q
        seqs['seq_' + src_entity + '_' + dst_entity] = q
        if p:
            data['params_' + name] = p
    for entity, n in counts.items():
        data['num_' + entity] = n
    for name, dst_entity in interleave_names:
        pattern = sample[name]
        max_len = {}
        for ent in pattern:
            if ent not in max_len:
                max_len[ent] = max(seqs['seq_' + ent + '_' + dst_entity]) + 1
        for ent, pos in interleave_indices(pattern, max_len).items():
            data['indices_' + ent + '_to_' + dst_entity] = pos
    return (data, output) if training else data


def read_dataset(directory: str, shuffle: bool = False, seed: Optional[int] = None) -> Iterator[dict]:
    """Yield raw samples from every ``*.tar.gz`` (each holding ``data.json``) of a directory.

    The file list is sorted before it is shuffled, and ``seed`` makes the shuffle reproducible: every rank of a
    data-parallel run must walk the files in the SAME order (it keeps a slice of each global batch)."""
    files = sorted(glob.glob(str(directory) + '/*.tar.gz'))
    if shuffle:
        (random.Random(seed) if seed is not None else random).shuffle(files)
    for path in files:
        with tarfile.open(path, 'r:gz') as tar:
            try:
                fh = tar.extractfile('data.json')
            except KeyError:
                raise DatasetError('The file data.json was not found in ' + path)
            for sample in json.load(fh):
                yield sample


def generator(directory, feature_names, output_name, adj_names, interleave_names, additional_input,
              training, shuffle=False):
    """Same positional signature as the reference generator (strings instead of bytes)."""
    for sample in read_dataset(directory, shuffle):
        yield sample_to_tensors(sample, feature_names, output_name, adj_names, interleave_names,
                                additional_input, training)


def find_dataset_dimensions(path: str) -> Dict[str, int]:
    """Per-key dimensions sniffed from the first sample (framework_operations.py:50-91)."""
    files = glob.glob(str(path) + '/*.tar.gz')
    if not files:
        raise DatasetError('No *.tar.gz file was found in ' + str(path))
    with tarfile.open(files[0], 'r:gz') as tar:
        sample = json.load(tar.extractfile('data.json'))[0]
    return sample_dimensions(sample)


def sample_dimensions(sample: dict) -> Dict[str, int]:
    dims: Dict[str, int] = {}
    for k, v in sample.items():
        if not isinstance(v, dict):
            dims[k] = len(v[0]) if isinstance(v, list) and v and isinstance(v[0], list) else 1
        elif v:
            first = v[next(iter(v))]
            if not isinstance(first[0], str) and isinstance(first[0], list):
                dims[k] = len(first[0][1])
            else:
                dims[k] = 0
    return dims
