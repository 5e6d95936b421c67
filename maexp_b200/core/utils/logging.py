"""Stdlib logger, as core/utils/logging.py:9-20 of the reference."""
import logging


def initialize_logging(name):
    logger = logging.getLogger(name)
    if not logging.getLogger().handlers:
        logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    return logger
