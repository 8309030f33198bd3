"""ignnition_b200: B200-native message-passing engine, drop-in for IGNNITION's generated model.

Public surface (same names as the reference, ``readme.md:77-95``): ``create_model``,
``train_and_evaluate``, ``predict``, ``debug``; plus ``Engine`` (the ``ComnetModel`` equivalent)
and ``ModelDescription`` (the ``Model_information`` equivalent).
"""

from .model_description import ModelDescription, Model_information, ModelDescriptionError  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not need torch / CUDA (host tools, oracle scripts)
    if name in ("Engine", "DeviceGraph"):
        from . import engine
        return getattr(engine, name)
    if name in ("create_model", "train_and_evaluate", "predict", "debug", "evaluate"):
        from . import framework_operations
        return getattr(framework_operations, name)
    raise AttributeError(name)
