"""The "MMSBM" logger: one file sink (``mmsbm.log`` in the working directory) and one console
sink, both at the requested level.  Calling ``setup_logger`` again replaces the sinks instead
of stacking them, and records never travel up to the root logger -- the observable behaviour
of the reference's src/logger.py:3-25."""
import logging

_LINE = "%(asctime)s %(levelname)s %(message)s"


def _drop_sinks(log):
    while log.handlers:
        sink = log.handlers[0]
        sink.close()
        log.removeHandler(sink)


def setup_logger(name, log_file="mmsbm.log", level=logging.INFO):
    log = logging.getLogger(name)
    _drop_sinks(log)
    layout = logging.Formatter(_LINE)
    sinks = [logging.FileHandler(log_file, encoding="utf-8"), logging.StreamHandler()]
    for sink in sinks:
        sink.setFormatter(layout)
        sink.setLevel(level)
        log.addHandler(sink)
    log.setLevel(level)
    log.propagate = False
    return log
