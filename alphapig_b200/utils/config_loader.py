"""Counterpart of reference ``utils/config_loader.py``: ``load_config(path)`` reads a YAML file into a dict and
``config_`` is the training configuration the reference loads at import time from ``../conf/train_config.yaml``
(:16-17).  Here ``config_`` comes from ``$ALPHAPIG_CONF`` or ``./conf/train_config.yaml`` when one exists, else from
the reference's committed defaults (``alphapig_b200.train_mxnet.DEFAULT_CONF``) with a console-only
``train_logging`` section - the reference's own section writes ``./logs/*.log`` files, which is a deployment
choice, not behaviour of the hot path."""
import os

import yaml


def load_config(data_path):
    with open(data_path, 'r') as f:
        return yaml.safe_load(f)


def _default():
    path = os.environ.get('ALPHAPIG_CONF') or os.path.join(os.getcwd(), 'conf', 'train_config.yaml')
    if os.path.isfile(path):
        conf = load_config(path)
        log = conf.get('train_logging') or {}
        # file handlers need their directory (the reference ships ./logs/ in the repo)
        for h in (log.get('handlers') or {}).values():
            fn = h.get('filename')
            if fn and os.path.dirname(fn):
                try:
                    os.makedirs(os.path.dirname(fn), exist_ok=True)
                except OSError:
                    pass
        return conf
    from ..train_mxnet import DEFAULT_CONF
    conf = dict(DEFAULT_CONF)
    conf['train_logging'] = {'version': 1, 'disable_existing_loggers': False}
    return conf


config_ = _default()
