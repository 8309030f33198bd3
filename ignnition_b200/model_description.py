"""Host-side reader for IGNNITION's ``model_description.json``.

The user surface is kept verbatim: the same JSON the reference validates and turns into a
``Model_information`` object tree (reference ``code/utils/json_operations.py:128-149``) is read
here into a small set of plain dataclasses.  Only what the message-passing hot path needs is kept
(entities, stages, sources, aggregation/update keywords, readout stack, learning options); the
getter names of the reference object are preserved so code written against it keeps working
(``json_operations.py:384-475``).

Nothing in this file touches the GPU.
"""

from __future__ import annotations

import copy
import json
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

AGGREGATIONS = ("sum", "ordered", "attention", "concat", "interleave", "convolution",
                # north-star extensions (no reference counterpart, defined by the oracle)
                "mean", "max")
MESSAGE_OPS = ("neural_network", "direct_assignation")
UPDATE_TYPES = ("neural_network", "recurrent_neural_network")
READOUT_TYPES = ("predict", "pooling", "product", "neural_network", "extend_adjacencies")


class ModelDescriptionError(ValueError):
    """Raised (with the reference's ``IGNNITION:`` prefix) when the JSON is malformed."""

    def __init__(self, msg: str):
        super().__init__("IGNNITION: " + msg)


@dataclass
class Feature:
    # reference: auxilary_classes.py:28-60
    name: str
    size: int = 1
    normalization: str = "None"


@dataclass
class Entity:
    # reference: auxilary_classes.py:62-127
    name: str
    hidden_state_dimension: int
    features: List[Feature] = field(default_factory=list)

    def get_entity_total_feature_size(self) -> int:
        return sum(f.size for f in self.features)

    def get_features_names(self) -> List[str]:
        return [f.name for f in self.features]


@dataclass
class DenseLayer:
    """One Keras-style layer of a feed-forward stack (auxilary_classes.py:800-866)."""
    type_layer: str
    name: str
    units: Optional[int] = None
    activation: Optional[str] = None
    kernel_regularizer: float = 0.0       # l2(lambda) = lambda * sum(w^2)
    use_bias: bool = True
    extra: Dict[str, Any] = field(default_factory=dict)


@dataclass
class FeedForward:
    """A Dense stack (auxilary_classes.py:868-975).  ``dropout``: the JSON holds Dropout layers with a non-zero rate
    (identity at inference, so they are not kept as layers; training through them is stochastic and not built)."""
    layers: List[DenseLayer]
    dropout: bool = False


@dataclass
class MessageOp:
    # reference: Operation / Apply_nn, auxilary_classes.py:163-205
    type: str                              # 'direct_assignation' | 'feed_forward_nn'
    input: List[str] = field(default_factory=list)
    output_name: str = "None"
    model: Optional[FeedForward] = None


@dataclass
class SourceEntity:
    # reference: Source_Entity, auxilary_classes.py:640-698
    name: str
    adj_vector: str
    message_formation: List[MessageOp]
    extra_parameters: int = 0

    def get_instance_info(self, dst_name: str) -> List[str]:
        return [self.adj_vector, self.name, dst_name, str(self.extra_parameters > 0)]


@dataclass
class Aggregation:
    # reference: auxilary_classes.py:229-456
    type: str
    interleave_definition: Optional[str] = None
    concat_axis: Optional[int] = None
    activation_function: str = "relu"      # convolution only

    @property
    def combination_definition(self):       # reference attribute name (Interleave_aggr)
        return self.interleave_definition


@dataclass
class Update:
    # reference: Apply_rnn / Apply_nn + Recurrent_Cell, auxilary_classes.py:186-227, 700-796
    type: str                              # 'recurrent_nn' | 'feed_forward_nn'
    recurrent_type: Optional[str] = None   # 'GRU'
    cell_parameters: Dict[str, Any] = field(default_factory=dict)
    model: Optional[FeedForward] = None


@dataclass
class MessagePassing:
    # reference: Message_Passing, auxilary_classes.py:458-638
    destination_entity: str
    source_entities: List[SourceEntity]
    aggregation: Aggregation
    update: Update

    def get_instance_info(self):
        return [s.get_instance_info(self.destination_entity) for s in self.source_entities]


@dataclass
class ReadoutOp:
    # reference: Readout_operation and subclasses, auxilary_classes.py:1033-1265
    type: str
    input: List[str]
    output_name: Any = None
    architecture: Optional[FeedForward] = None
    label: Optional[str] = None
    label_normalization: Optional[str] = None
    label_denormalization: Optional[str] = None
    type_pooling: Optional[str] = None
    type_product: Optional[str] = None
    adj_list: Optional[str] = None


def _require(d: dict, keys, where: str):
    for k in keys:
        if k not in d:
            raise ModelDescriptionError("'%s' is a required property of %s" % (k, where))


def _parse_layers(arch: list, role: str) -> FeedForward:
    layers = []
    dropout = False
    for i, l in enumerate(arch):
        l = dict(l)
        _require(l, ["type_layer"], "a layer of the %s neural network" % role)
        t = l.pop("type_layer")
        if t == "Dropout":                 # tf.keras.layers.Dropout: the identity outside training
            dropout = dropout or float(l.get("rate", 0.0)) > 0.0
            continue
        # default layer name: auxilary_classes.py:909-910
        name = l.pop("name", "layer_%d_%s_%s" % (i, t, role))
        act = l.pop("activation", None)
        if act == "None":
            act = None
        reg = float(l.pop("kernel_regularizer", 0.0))
        units = l.pop("units", None)
        use_bias = l.pop("use_bias", True)
        if isinstance(use_bias, str):
            use_bias = use_bias == "True"
        layers.append(DenseLayer(t, name, None if units is None else int(units), act, reg,
                                 bool(use_bias), l))
    return FeedForward(layers, dropout)


class ModelDescription:
    """Equivalent of the reference's ``Model_information`` (json_operations.py:29-476)."""

    def __init__(self, path_or_dict, dimensions: Optional[Dict[str, int]] = None):
        if isinstance(path_or_dict, dict):
            data = copy.deepcopy(path_or_dict)
        else:
            with open(path_or_dict) as fh:
                data = json.load(fh)
        dimensions = dict(dimensions or {})
        self._validate(data)

        self.nn_architectures = {m["nn_name"]: m for m in data["neural_networks"]}
        self.entities = [self._entity(e, dimensions) for e in data["entities"]]
        self.iterations_mp = int(data["message_passing"]["num_iterations"])
        self.mp_instances = [[st["stage_name"], [self._mp(m, dimensions) for m in st["stage_mp"]]]
                             for st in data["message_passing"]["stages"]]
        self.readout_op = [self._readout(op) for op in data["readout"]]
        self.training_op = {"loss": data["learning_options"]["loss"],
                            "optimizer": copy.deepcopy(data["learning_options"]["optimizer"])}
        # json_operations.py:370-382: entity hidden sizes merged with the dataset dimensions
        self.input_dim = {**{e.name: e.hidden_state_dimension for e in self.entities}, **dimensions}

    # ------------------------------------------------------------------ validation
    def _validate(self, data: dict):
        """Structural + cross-reference checks (json_operations.py:138-139, 184-245)."""
        _require(data, ["entities", "message_passing", "readout", "neural_networks",
                        "learning_options"], "the model description")
        for e in data["entities"]:
            _require(e, ["name", "hidden_state_dimension", "features"], "an entity")
            for f in e["features"]:
                _require(f, ["name"], "a feature")
        mp = data["message_passing"]
        _require(mp, ["num_iterations", "stages"], "message_passing")
        _require(data["learning_options"], ["loss", "optimizer"], "learning_options")
        entity_names = [e["name"] for e in data["entities"]]
        nn_names = []
        for n in data["neural_networks"]:
            _require(n, ["nn_name", "nn_type"], "a neural network")
            if n["nn_type"] not in ("feed_forward", "recurrent_neural_network"):
                raise ModelDescriptionError("'%s' is not a valid nn_type" % n["nn_type"])
            if n["nn_type"] == "feed_forward":
                _require(n, ["nn_architecture"], "a feed_forward neural network")
            else:
                _require(n, ["recurrent_type"], "a recurrent neural network")
                if n["recurrent_type"] not in ("GRU", "LSTM"):
                    raise ModelDescriptionError("'%s' is not a valid recurrent_type" % n["recurrent_type"])
            nn_names.append(n["nn_name"])
        called, inputs, outputs = [], [], ["hs_source", "hs_dest", "edge_params"]
        for st in mp["stages"]:
            _require(st, ["stage_name", "stage_mp"], "a stage")
            for m in st["stage_mp"]:
                _require(m, ["source_entities", "destination_entity", "aggregation", "update"],
                         "a message passing")
                if m["destination_entity"] not in entity_names:
                    raise ModelDescriptionError(
                        "The destination entity " + m["destination_entity"] +
                        " was used in a message passing. However, there is no such entity. \n"
                        " Please check the spelling or define a new entity.")
                agg = m["aggregation"]
                _require(agg, ["type"], "an aggregation")
                if agg["type"] not in AGGREGATIONS:
                    raise ModelDescriptionError("'%s' is not a valid aggregation" % agg["type"])
                if agg["type"] == "interleave":
                    _require(agg, ["interleave_definition"], "an interleave aggregation")
                if agg["type"] == "concat":
                    _require(agg, ["concat_axis"], "a concat aggregation")
                    if int(agg["concat_axis"]) not in (1, 2):
                        raise ModelDescriptionError("concat_axis must be 1 or 2")
                up = m["update"]
                _require(up, ["type"], "an update")
                if up["type"] not in UPDATE_TYPES:
                    raise ModelDescriptionError("'%s' is not a valid update type" % up["type"])
                _require(up, ["nn_name"], "an update")
                called.append(up["nn_name"])
                for s in m["source_entities"]:
                    _require(s, ["name", "adj_vector", "message"], "a source entity")
                    if s["name"] not in entity_names:
                        raise ModelDescriptionError(
                            "The source entity " + s["name"] +
                            " was used in a message passing. However, there is no such entity. \n"
                            " Please check the spelling or define a new entity.")
                    for op in s["message"]:
                        _require(op, ["type"], "a message operation")
                        if op["type"] not in MESSAGE_OPS:
                            raise ModelDescriptionError("'%s' is not a valid message operation" % op["type"])
                        if op["type"] == "neural_network":
                            _require(op, ["nn_name", "input"], "a neural_network message operation")
                            called.append(op["nn_name"])
                            inputs += op["input"]
                        if "output_name" in op:
                            outputs.append(op["output_name"])
        for op in data["readout"]:
            _require(op, ["type", "input"], "a readout operation")
            if op["type"] not in READOUT_TYPES:
                raise ModelDescriptionError("'%s' is not a valid readout operation" % op["type"])
            if op["type"] == "predict":
                _require(op, ["nn_name", "label"], "a predict operation")
            if op["type"] == "neural_network":
                _require(op, ["nn_name", "output_name"], "a readout neural_network")
            if op["type"] == "pooling":
                _require(op, ["type_pooling", "output_name"], "a pooling operation")
            if op["type"] == "product":
                _require(op, ["type_product", "output_name"], "a product operation")
            if op["type"] == "extend_adjacencies":
                _require(op, ["adj_list", "output_name_src", "output_name_dst"], "extend_adjacencies")
            if op["type"] in ("predict", "neural_network"):
                called.append(op["nn_name"])
        for name in called:
            if name not in nn_names:
                raise ModelDescriptionError(
                    "The name " + name + " is used as a reference to a neural network (nn_name), "
                    "even though the neural network was not defined. \n Please make sure the name "
                    "is correctly spelled or define a neural network named " + name)
        for i in inputs:
            if i not in outputs:
                raise ModelDescriptionError(
                    "The name " + i + " was used as input of a message creation operation even "
                    "though it wasn't the output of one.")

    # ------------------------------------------------------------------ builders
    def _entity(self, e: dict, dims: Dict[str, int]) -> Entity:
        feats = []
        for f in e.get("features", []):
            # json_operations.py:162-176: the dataset decides the feature size
            size = int(dims.get(f["name"], f.get("size", 1)))
            feats.append(Feature(f["name"], size, str(f.get("normalization", "None"))))
        return Entity(e["name"], int(e["hidden_state_dimension"]), feats)

    def _ff(self, nn_name: str, role: str) -> FeedForward:
        info = self.nn_architectures[nn_name]
        if info["nn_type"] != "feed_forward":
            raise ModelDescriptionError("The neural network " + nn_name + " is not a feed_forward model")
        return _parse_layers(copy.deepcopy(info["nn_architecture"]), role)

    def _mp(self, m: dict, dims: Dict[str, int]) -> MessagePassing:
        # json_operations.py:270-300 (__add_nn_architecture) + Message_Passing ctor
        sources = []
        for s in m["source_entities"]:
            ops = []
            for k, op in enumerate(s["message"]):
                if op["type"] == "neural_network":
                    ops.append(MessageOp("feed_forward_nn", list(op["input"]),
                                         op.get("output_name", "None"),
                                         self._ff(op["nn_name"], "message_creation_%d" % k)))
                else:
                    ops.append(MessageOp("direct_assignation"))
            sources.append(SourceEntity(s["name"], s["adj_vector"], ops,
                                        int(dims.get(s["adj_vector"], 0))))
        a = m["aggregation"]
        agg = Aggregation(a["type"], a.get("interleave_definition"),
                          int(a["concat_axis"]) if "concat_axis" in a else None,
                          a.get("activation_function", "relu"))
        u = m["update"]
        if u["type"] == "recurrent_neural_network":
            arch = self.nn_architectures[u["nn_name"]]
            if arch["nn_type"] != "recurrent_neural_network":
                raise ModelDescriptionError("The neural network " + u["nn_name"] + " is not recurrent")
            params = {k: v for k, v in arch.items() if k not in ("nn_name", "nn_type", "recurrent_type")}
            upd = Update("recurrent_nn", arch["recurrent_type"], params)
        else:
            upd = Update("feed_forward_nn", model=self._ff(u["nn_name"], "update"))
        return MessagePassing(m["destination_entity"], sources, agg, upd)

    def _readout(self, op: dict) -> ReadoutOp:
        t = op["type"]
        r = ReadoutOp(t, list(op["input"]))
        if t in ("predict", "neural_network"):
            r.architecture = self._ff(op["nn_name"], "readout")
        if t == "predict":
            r.label = op["label"]
            r.label_normalization = op.get("label_normalization")
            r.label_denormalization = op.get("label_denormalization")
        elif t == "neural_network":
            r.output_name = op.get("output_name", "None")
        elif t == "pooling":
            r.type_pooling, r.output_name = op["type_pooling"], op["output_name"]
        elif t == "product":
            r.type_product, r.output_name = op["type_product"], op["output_name"]
        elif t == "extend_adjacencies":
            r.adj_list = op["adj_list"]
            r.output_name = [op["output_name_src"], op["output_name_dst"]]
        return r

    # ------------------------------------------------------------------ reference getters
    def get_input_dimensions(self): return self.input_dim
    def get_entities(self): return self.entities
    def get_mp_iterations(self): return self.iterations_mp
    def get_mp_instances(self): return self.mp_instances
    def get_optimizer(self): return copy.deepcopy(self.training_op["optimizer"])
    def get_loss(self): return self.training_op["loss"]
    def get_readout_operations(self): return self.readout_op

    def get_interleave_sources(self):
        return [[s.name, mp.destination_entity] for _, mps in self.mp_instances for mp in mps
                if mp.aggregation.type == "interleave" for s in mp.source_entities]

    def get_interleave_tensors(self):
        return [[mp.aggregation.interleave_definition, mp.destination_entity]
                for _, mps in self.mp_instances for mp in mps if mp.aggregation.type == "interleave"]

    def get_output_info(self):
        p = [o for o in self.readout_op if o.type == "predict"][0]
        return p.label, p.label_normalization, p.label_denormalization

    def get_all_features(self):
        return [f for e in self.entities for f in e.features]

    def get_adjecency_info(self):          # (sic) reference spelling, json_operations.py:451
        return [info for _, mps in self.mp_instances for mp in mps for info in mp.get_instance_info()]

    def get_additional_input_names(self):
        outs, ins = set(), set()
        for r in self.readout_op:
            if r.type == "extend_adjacencies":
                outs.update(r.output_name)
            elif r.type != "predict":
                outs.add(r.output_name)
            ins.update(r.input)
        outs.update(e.name for e in self.entities)
        return sorted(ins.difference(outs))


# reference class name, kept as an alias for drop-in use
Model_information = ModelDescription
