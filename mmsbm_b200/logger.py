"""Named logger with a file and a console handler, re-configured on every call
(behaviour of the reference's src/logger.py:3-25)."""
import logging


def setup_logger(name, log_file="mmsbm.log", level=logging.INFO):
    log = logging.getLogger(name)
    log.setLevel(level)
    for old in list(log.handlers):
        old.close()
        log.removeHandler(old)
    fmt = logging.Formatter("%(asctime)s %(levelname)s %(message)s")
    for handler in (logging.FileHandler(log_file, encoding="utf-8"), logging.StreamHandler()):
        handler.setLevel(level)
        handler.setFormatter(fmt)
        log.addHandler(handler)
    log.propagate = False
    return log
