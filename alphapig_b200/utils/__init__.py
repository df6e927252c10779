"""Counterparts of the reference ``utils`` package (utils/__init__.py:2-3 imports ``config_loader`` and
``sgf_dataIter``): the SGF reader the training path needs, the YAML config loader, and an inert ``send_email``."""
from . import sgf_dataIter  # noqa: F401


def __getattr__(name):
    # config_loader imports train_mxnet's defaults lazily (train_mxnet imports this package)
    if name in ("config_loader", "send_email"):
        import importlib
        return importlib.import_module("%s.%s" % (__name__, name))
    raise AttributeError(name)
