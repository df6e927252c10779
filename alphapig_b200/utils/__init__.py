"""Counterparts of the reference ``utils`` package that the training path needs (SGF reader only;
e-mail and YAML/logging config are out of scope)."""
from . import sgf_dataIter  # noqa: F401
