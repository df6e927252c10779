"""Counterpart of reference ``utils/send_email.py`` so that ``from utils import config_loader, send_email``
(train_mxnet.py:22) keeps importing.  Sending mail is outside this engine's scope (and the reference hard-codes
its author's account): ``send_mail`` logs the message and reports "not sent"."""
import logging

_logger = logging.getLogger(__name__)


def send_mail(message_title, message_text, pass_wd):
    _logger.info("send_mail (not sent, e-mail is out of scope): %s - %s", message_title, message_text)
    return False
