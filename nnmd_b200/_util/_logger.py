"""File logger used by `BPNN.fit` (reference: nnmd/_util/_logger.py:4-19): file handler, mode "w", no propagation."""
import logging


class Logger:
    @classmethod
    def get_logger(cls, logger_name, filename: str, level=logging.INFO) -> logging.Logger:
        log = logging.getLogger(logger_name)
        log.setLevel(level)
        log.addHandler(logging.FileHandler(filename, mode="w"))
        log.propagate = False
        return log
